set -x
python -m pytest tests/test_gpu_scene.py tests/test_gpu_fused.py -m gpu -x -q 2>&1 | tail -8
python tools/scene_probe.py 256
python - <<'PY'
import numpy as np, torch, sys
sys.path.insert(0,'.')
import tod_b200
from tests import synth
nb=256
sb=tod_b200.SceneBuilder(max_batch=nb)
depth=torch.from_numpy(np.tile(synth.depth_frames(8,seed=3),(nb//8,1,1)).astype(np.int16)).cuda()
target=torch.zeros_like(depth)
npx=640*480
o_map=torch.empty((nb,npx),dtype=torch.int32,device='cuda'); o_w=torch.empty((nb,npx,4),dtype=torch.float32,device='cuda'); o_c0=torch.empty_like(o_w); o_c1=torch.empty_like(o_w); o_b=torch.empty((nb,100,4),dtype=torch.float32,device='cuda')
st=torch.cuda.Stream()
def step(): sb.append_batch_device(depth.data_ptr(),target.data_ptr(),nb,o_map.data_ptr(),o_w.data_ptr(),o_c0.data_ptr(),o_c1.data_ptr(),o_b.data_ptr(),st.cuda_stream)
for _ in range(3): step()
torch.cuda.synchronize()
e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
with torch.cuda.stream(st):
    e0.record()
    for _ in range(5): step()
    e1.record()
torch.cuda.synchronize()
print("scene total ms/256 frames:", e0.elapsed_time(e1)/5)
PY
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/scene_launches_r2b.csv python tools/scene_probe.py 64 > gpurun_out/ncu_scene.log 2>&1
grep -E "land_kernel|stamp_|merge_kernel|weights_kernel|balls" gpurun_out/scene_launches_r2b.csv | awk -F'","' '{print substr($5,1,40), $NF}' | tail -5
