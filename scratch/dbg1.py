import sys, tempfile, numpy as np
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import util
from oracle import host as OH, refcl as OR
rt = util.product()
COLS,ROWS,RPP=96,64,4
o_scene,p_scene = util.make_scene_pair(tempfile.mkdtemp(), COLS, ROWS, mesh_uv=(32,16), mesh_nslabs=10)
lib = OR.load_best()
total=COLS*ROWS*RPP
seeds0=OR.make_seeds(total,11)
prep=OR.prepare_a10(o_scene,1)
cam=o_scene['camera'].toFloat32Array()
print('cam equal', np.array_equal(cam, p_scene['camera'].toFloat32Array()))
for i,(a,b) in enumerate(zip(o_scene['lights'],p_scene['lights'])):
    print('light',i,[np.array_equal(getattr(a,f)(),getattr(b,f)()) for f in ('toShadowInfo','toSceneRenderInfo','toLightRenderInfo')])
print('mat', np.array_equal(OH.splitMaterialData(o_scene), rt.splitMaterialData(p_scene)))
for depth in (0,1,5):
  for tile in (total, COLS*8*RPP):
    st=OR.A10State(total,seeds0); lib.a10_initAcu(st.acu,total)
    OR.a10_execute_render(lib,st,prep,cam,COLS,ROWS,RPP,o_scene['focal_length'],o_scene['lens_diameter'],depth=depth)
    r=rt.Renderer(p_scene,COLS,ROWS,RPP,mode=1,tile_slots=tile,depth=depth); r.preRender(seeds0)
    pix=r.executeRender(); acc=r.accum()
    ref=st.acu.reshape(COLS*ROWS,RPP,4).sum(1)
    sd=r.seeds()
    print('depth',depth,'tile',tile,'maxdiff',np.abs(acc-ref).max(),'seeds eq',np.array_equal(sd,st.seeds), 'nbad', np.count_nonzero(np.abs(acc-ref).max(1)>1e-5), r.stats())
    r.postRender()
