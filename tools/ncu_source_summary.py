"""Summarise an .ncu-rep per CUDA source file / line: stall samples, warp instructions, threads per instruction.

usage: python tools/ncu_source_summary.py report.ncu-rep [top_n]
Uses NVIDIA's ncu_report module (SASS pc -> source line correlation from -lineinfo).
"""
import sys, collections, os
sys.path.insert(0, "/opt/nvidia/nsight-compute/2025.2.1/extras/python")
import ncu_report

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
ctx = ncu_report.load_report(rep)
for ri in range(ctx.num_ranges()):
    rng = ctx.range_by_idx(ri)
    for ai in range(rng.num_actions()):
        act = rng.action_by_idx(ai)
        print("== kernel", act.name(), "duration_ms", act.metric_by_name("gpu__time_duration.sum").as_double() / 1e6)
        def inst_metric(name):
            m = act.metric_by_name(name)
            if m is None: return {}
            pcs = m.correlation_ids()
            return {pcs.as_uint64(i): m.as_uint64(i) if m.kind() != ncu_report.IMetric.ValueKind_DOUBLE else m.as_double(i) for i in range(m.num_instances())}
        samples = inst_metric("smsp__pcsamp_sample_buffer_full") and {}
        samples = inst_metric("smsp__pcsamp_warps_issue_stalled_all") if False else {}
        # total samples per pc = sum of all stall reasons
        stall_names = [n for n in act.metric_names() if n.startswith("smsp__pcsamp_warps_issue_stalled_") and not n.endswith("_not_issued")]
        per_pc = collections.defaultdict(lambda: collections.defaultdict(float))
        for n in stall_names:
            short = n[len("smsp__pcsamp_warps_issue_stalled_"):]
            for pc, v in inst_metric(n).items():
                per_pc[pc][short] += v
        inst = inst_metric("inst_executed")
        thr = inst_metric("thread_inst_executed")
        lines = collections.defaultdict(lambda: [0.0, 0.0, 0.0, collections.defaultdict(float)])
        for pc in set(list(per_pc.keys()) + list(inst.keys())):
            si = act.source_info(pc)
            key = (os.path.basename(si.file_name()), si.line()) if si is not None else ("?", 0)
            L = lines[key]
            for k, v in per_pc.get(pc, {}).items():
                L[0] += v; L[3][k] += v
            L[1] += inst.get(pc, 0); L[2] += thr.get(pc, 0)
        ts = sum(v[0] for v in lines.values()); ti = sum(v[1] for v in lines.values())
        print(f"total samples {ts:.0f} warp inst {ti:.0f}")
        byfile = collections.defaultdict(lambda: [0.0, 0.0, 0.0])
        for (f, ln), v in lines.items():
            byfile[f][0] += v[0]; byfile[f][1] += v[1]; byfile[f][2] += v[2]
        for f, (s, ins, t) in sorted(byfile.items(), key=lambda kv: -kv[1][0]):
            print(f"  {f:28s} samples {100*s/max(ts,1):5.1f}% inst {100*ins/max(ti,1):5.1f}% thr {t/max(ins,1):4.1f}")
        srccache = {}
        def srcline(f, ln):
            for d in ("homework-18-graphics-raytracer_b200/csrc",):
                p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), d, f)
                if os.path.exists(p):
                    if p not in srccache: srccache[p] = open(p).read().split("\n")
                    if 0 < ln <= len(srccache[p]): return srccache[p][ln - 1].strip()[:80]
            return ""
        for (f, ln), v in sorted(lines.items(), key=lambda kv: -kv[1][0])[:top]:
            top3 = sorted(v[3].items(), key=lambda kv: -kv[1])[:3]
            top3 = [(k, int(100 * x / max(v[0], 1))) for k, x in top3]
            print(f"  {f}:{ln:4d} samples {100*v[0]/max(ts,1):4.1f}% inst {100*v[1]/max(ti,1):4.1f}% thr {v[2]/max(v[1],1):4.1f} {top3} | {srcline(f, ln)}")
        # optional: aggregate by named line ranges, e.g.  --ranges rt_cast.cuh:filter=194-211,241-258;confirm=82-100,259-274
        for arg in sys.argv[3:]:
            if not arg.startswith("--ranges="): continue
            fname, spec = arg[len("--ranges="):].split(":", 1)
            print(f"  ranges in {fname}:")
            for part in spec.split(";"):
                name, rs = part.split("=")
                s = ins = t = 0.0
                for r in rs.split(","):
                    a, b = (int(x) for x in r.split("-"))
                    for (f, ln), v in lines.items():
                        if f == fname and a <= ln <= b:
                            s += v[0]; ins += v[1]; t += v[2]
                print(f"    {name:12s} samples {100*s/max(ts,1):5.1f}% inst {100*ins/max(ti,1):5.1f}% thr {t/max(ins,1):4.1f}")
