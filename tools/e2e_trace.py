"""Where the end-to-end step of bench.py spends its time beyond the kernel (diagnostic)."""
import sys, time
sys.path.insert(0, ".")
import torch
import go_raytracer_b200 as g
s, cfg = g.builtin_scene(6, width=1024, spp=4096)
cam = g.derive_camera(cfg)
nval = cam.width * cam.height * 3
dev = torch.device("cuda", 0)
acc = torch.zeros(nval, dtype=torch.float32, device=dev)
rgb8 = torch.zeros(nval, dtype=torch.uint8, device=dev)
h_sum = torch.zeros(nval, dtype=torch.float32).pin_memory()
h_rgb8 = torch.zeros(nval, dtype=torch.uint8).pin_memory()
h_zero = torch.zeros(nval, dtype=torch.float32).pin_memory()
stream = torch.cuda.current_stream()
def T(): torch.cuda.synchronize(); return time.perf_counter()
for rep in range(3):
    t0 = T(); sc = g.DeviceScene(s, 0); t1 = T()
    acc.copy_(h_zero, non_blocking=True); t2 = T()
    sc.render_device(cam, acc.data_ptr(), stream.cuda_stream); t3 = T()
    sc.tonemap_device(acc.data_ptr(), rgb8.data_ptr(), nval, 1.0 / 4096, stream.cuda_stream); t4 = T()
    h_sum.copy_(acc, non_blocking=True); h_rgb8.copy_(rgb8, non_blocking=True); t5 = T()
    sc.close(); t6 = T()
    print("upload %.2f ms  h2d %.2f  render %.2f  tonemap %.2f  d2h %.2f  free %.2f" % tuple(1e3 * x for x in (t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4, t6 - t5)))
