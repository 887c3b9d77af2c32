import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np, fastlivo_loader
flb=fastlivo_loader.load(); po=fastlivo_loader.oracle()
seq=flb.synth.make_visual_sequence("T0",4)
cam=seq["cam"]
ovm=po.VMap(cam,grid_size=16,outlier_threshold=300.0)
h=flb.Handle(device=0); h.camera_set(cam); h.image_upload(seq["frames"][0]["image"]); h.vmap_reset(seq,grid_size=16,outlier_threshold=300.0)
for fr in seq["frames"]:
    ovm.select(fr["image"],fr["Rcw"],fr["Pcw"],fr["pg_down"]); omv0=ovm.map_value()
    h.image_upload(fr["image"]); h.vmap_select(fr["Rcw"],fr["Pcw"],fr["pg_down"]); gmv0=h.vmap_map_value(320)
    print("frame",fr["frame_id"],"mv0 equal",(omv0==gmv0).all())
    omv,owin=po.visual_candidates(cam,fr["Rcw"],fr["Pcw"],fr["image"],fr["pg"],16,40,omv0)
    gmv_old,gwin_old=h.visual_candidates(fr["Rcw"],fr["Pcw"],fr["pg"],16,40,gmv0)
    print("  old API winners equal",(owin==gwin_old).all(), (owin>=0).sum(), (gwin_old>=0).sum())
    ovm.grow(fr["image"],fr["Rcw"],fr["Pcw"],fr["pg"],fr["frame_id"])
    h.vmap_grow(fr["Rcw"],fr["Pcw"],fr["pg"],fr["frame_id"]); c=h.vmap_counts(); gmv=h.vmap_map_value(320)
    bad=np.nonzero(gmv!=ovm.map_value())[0]
    print("  grow last_added",c["last_added"],"oracle winners",(owin>=0).sum(),"bad cells",bad, gmv[bad], ovm.map_value()[bad], "owin",owin[bad])
    for b in bad:
        i=owin[b]; p=fr["pg"][i].astype(np.float64); pf=fr["Rcw"]@p+fr["Pcw"]
        print("    cell",b,"pt idx",i,p,"pc",cam["fx"]*pf[0]/pf[2]+cam["cx"],cam["fy"]*pf[1]/pf[2]+cam["cy"])
    ovm.add_observations(fr["image"],fr["Rcw"],fr["Pcw"],fr["frame_id"]); h.vmap_add_observations(fr["Rcw"],fr["Pcw"],fr["frame_id"]); print("  obs", h.vmap_counts())
# colorize
seq=flb.synth.make_visual_sequence("T1",1); fr=seq["frames"][0]
rng=np.random.default_rng(2)
bgr=np.stack([fr["image"],np.roll(fr["image"],3,1),255-fr["image"]],-1)^rng.integers(0,8,fr["image"].shape+(3,),dtype=np.uint8)
pts=np.concatenate([fr["pg"],-fr["pg"][:50],fr["pg"][:50]*np.float32(40.0)])
h2=flb.Handle(device=0); h2.camera_set(seq["cam"])
rgb,val=h2.colorize(fr["Rcw"],fr["Pcw"],bgr,pts); orgb,oval=po.colorize(seq["cam"],fr["Rcw"],fr["Pcw"],bgr,pts)
d=np.nonzero((rgb!=orgb).any(1))[0]
print("colorize: differing points",len(d),"of",len(pts)); print(rgb[d[:8]],orgb[d[:8]], val[d[:8]])
