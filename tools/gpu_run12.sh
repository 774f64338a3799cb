python bench.py --workload cfg3 --steps 1 --warmup 1 --no-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('cfg3 screen', d['value'], d['ms_per_step'], d['roofline']['match_ms_per_view'], d['roofline']['build_ms_per_view'])"
SR_MATCH_SCREEN=0 python bench.py --workload cfg3 --steps 1 --warmup 1 --no-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('cfg3 fp64', d['value'], d['ms_per_step'], d['roofline']['match_ms_per_view'], d['roofline']['build_ms_per_view'])"
python - <<'PY'
import time, numpy as np
from stereoreconstruction_b200 import capi, scenes, types as T
import os
w,h,V,D=1920,1080,8,256
cams=scenes.arc_cameras(V,w,h)
surf=scenes.HeightField(z0=0.0,amp=25.0,lx=90.0,ly=70.0)
for flag in ("1","0"):
    os.environ["SR_MATCH_SCREEN"]=flag
    c=capi.Context(0)
    c.set_views(cams,[np.zeros((h,w,4),np.uint8)]*V,None)
    P=T.default_params(False,350.0,650.0,D); c.set_params(P)   # two-view defaults: r=5 geodesic NCC
    if flag=="1":
        rays={v:c.unproject_grid(v) for v in (3,4)}
        imgs=scenes.render_views(V,lambda v: rays[v] if v in rays else rays[3],surf,4321,3.5*500.0/cams[0].K[0])
    c.set_views(cams,imgs,None); c.set_params(P)
    for it in range(2):
        c.synchronize(); t0=time.perf_counter(); c.run_view(3,[4]); c.synchronize(); t1=time.perf_counter()
    print('two-view r=5 geodesic 1080p x 256 labels refractive, screen=%s: %.1f ms'%(flag,(t1-t0)*1e3))
    c.close()
PY
