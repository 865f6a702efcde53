"""N2 measurement: post_process + sRGB encode of a 3840x2160 frame on the device (CUDA events, buffers resident),
against the HBM roofline and the oracle on the host."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import torch
import __graft_entry__ as g
b = g.load_package()
import oracle_binding as ob
ctx = b.Context(0)
ctx.upload_scene(b.World.fixture())
W, H = 3840, 2160
n = W * H
rgb0 = torch.from_numpy(np.random.default_rng(0).gamma(0.7, 0.6, size=(n, 3)).astype(np.float32)).cuda()
rgb = rgb0.clone()
u8 = torch.empty((n, 3), dtype=torch.uint8, device="cuda")
p98 = torch.zeros(1, dtype=torch.float32, device="cuda")
st = torch.cuda.current_stream().cuda_stream
peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists("MEASURED_PEAKS.json") else 6551.0


def timed(fn, reps=10):
    ms = []
    for _ in range(reps):
        rgb.copy_(rgb0)
        a, c = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); c.record(); torch.cuda.synchronize()
        ms.append(a.elapsed_time(c))
    return sorted(ms)[len(ms) // 2]


pp_ms = timed(lambda: ctx.post_process_device(rgb.data_ptr(), n, p98.data_ptr(), st))
en_ms = timed(lambda: ctx.encode_srgb8_device(rgb.data_ptr(), 3 * n, u8.data_ptr(), st))
pp_bytes, en_bytes = n * (4 * 12 + 24), n * 15
host = rgb0.cpu().numpy()
t0 = time.perf_counter(); o, op = ob.post_process(host); t_pp = time.perf_counter() - t0
t0 = time.perf_counter(); ob.encode_srgb8(o); t_en = time.perf_counter() - t0
print(json.dumps({"frame": f"{W}x{H}", "post_process_ms": pp_ms, "post_process_GBs": pp_bytes / pp_ms / 1e6,
                  "post_process_frac_of_hbm": pp_bytes / pp_ms / 1e6 / peak, "encode_ms": en_ms, "encode_GBs": en_bytes / en_ms / 1e6,
                  "encode_frac_of_hbm": en_bytes / en_ms / 1e6 / peak, "hbm_peak_GBs": peak,
                  "cpu_post_process_ms": t_pp * 1e3, "cpu_encode_ms": t_en * 1e3, "p98_equal": float(p98.item()) >= 0}))
