python bench.py --workload cfg3 --steps 1 --warmup 1 --no-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('cfg3', d['value'], d['ms_per_step'], d['roofline']['match_ms_per_view'], d['roofline']['build_ms_per_view'])"
python - <<'PY'
import time, numpy as np, torch
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
for it in range(2):
    c.synchronize(); t0=time.perf_counter(); c.run_view_curve(3,[2,4,5]); c.synchronize(); t1=time.perf_counter()
    print('curve mode cfg4 view 3: %.1f ms'%((t1-t0)*1e3))
dc=c.depth(3).copy()
c.run_view(3,[2,4,5]); dl=c.depth(3)
both=(dc>0)&(dl>0)
print('labelled curve %.3f label %.3f; median |curve depth - label depth| = %.3f (label step %.3f)'%((dc>0).mean(),(dl>0).mean(),np.median(np.abs(dc[both]-dl[both])),300/255))
PY
