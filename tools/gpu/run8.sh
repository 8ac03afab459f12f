python -m pytest tests/test_gpu_scene.py tests/test_gpu_property.py::test_scene_random_frames -m gpu -x -q 2>&1 | tail -3
python tools/scene_probe.py 256
python - <<'PY'
import torch
x=torch.empty(1<<30,dtype=torch.float32,device='cuda')
for _ in range(2): x.zero_()
torch.cuda.synchronize()
e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): x.zero_()
e1.record(); torch.cuda.synchronize()
print("memset 4 GiB: %.1f GB/s" % (5*4.294967296/ (e0.elapsed_time(e1)*1e-3)))
y=torch.empty_like(x)
e0.record()
for _ in range(5): y.copy_(x)
e1.record(); torch.cuda.synchronize()
print("copy 4 GiB: %.1f GB/s (read+write)" % (5*2*4.294967296/ (e0.elapsed_time(e1)*1e-3)))
PY
