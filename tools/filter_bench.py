import sys
sys.path.insert(0,'.')
import __graft_entry__ as g
b = g.load_package()
ctx = b.Context(0)
info = ctx.device_info()
peak, mhz = ctx.measure_fp32_peak()
print("ffma loop TFLOP/s", peak, "eff MHz", mhz, info)
nominal = info['sm_count']*128*2*1.965e9/1e12
for bps in (2, 4, 8):
    for v in (0, 1, 2, 3, 4, 5, 6, 7):
        ms, pairs = ctx.filter_bench(v, bps, 256)
        tf = pairs*36/ (ms*1e-3)/1e12
        print(f"variant {v} blocks/SM {bps}: {ms:.3f} ms  {pairs/ms/1e6:.1f} Gpairs/s  algorithmic {tf:.2f} TFLOP/s = {100*tf/nominal:.1f}% of {nominal:.1f}")

names = ["FFMA a=a*b+c", "FFMA2 a2=a2*b2+c2", "FFMA2 bcast", "FFMA 3 live regs", "FMNMX", "MUFU.RCP", "3 FFMA2 : 1 FMNMX(+conv)", "FFMA2 3 live regs"]
for v in range(8):
    ms, ipc = ctx.pipe_bench(v)
    print(f"pipe {v} {names[v]:28s}: {ms:.3f} ms  {ipc:.3f} warp-inst/clk/SMSP (at 1965 MHz)")
