# Times sr_run_view_curve on one cfg4 reference view (run from the repo root: PYTHONPATH=. python tools/curve_time.py)
import time, numpy as np
from stereoreconstruction_b200 import capi, scenes, types as T
w,h,V,D=1920,1080,8,256
cams=scenes.arc_cameras(V,w,h)
c=capi.Context(0)
c.set_views(cams,[np.zeros((h,w,4),np.uint8)]*V,None)
P=T.default_params(True,350.0,650.0,D); c.set_params(P)
surf=scenes.HeightField(z0=0.0,amp=25.0,lx=90.0,ly=70.0)
rays={v:c.unproject_grid(v) for v in (2,3,4,5)}
imgs=scenes.render_views(V,lambda v: rays[v] if v in rays else rays[3],surf,4321,3.5*500.0/cams[0].K[0])
c.set_views(cams,imgs,None); c.set_params(P)
c.set_profiling(True)
for it in range(3):
    c.synchronize(); t0=time.perf_counter(); c.run_view_curve(3,[2,4,5]); c.synchronize(); t1=time.perf_counter()
    st=c.stage_ms()
    print('curve mode cfg4 view 3: %.1f ms (curve build %.1f, weights + match %.1f)'%((t1-t0)*1e3, st['build_ms'], st['match_ms']))
for it in range(2):
    c.synchronize(); t0=time.perf_counter(); c.run_view(3,[2,4,5]); c.synchronize(); t1=time.perf_counter()
    st=c.stage_ms()
    print('label mode cfg4 view 3: %.1f ms (build %.1f, weights + match %.1f)'%((t1-t0)*1e3, st['build_ms'], st['match_ms']))
