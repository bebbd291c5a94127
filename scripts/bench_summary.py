"""Print the key numbers of a bench.py JSON line (development aid)."""
import json
import sys

d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
r = d["roofline"]
print(f"value {d['value'] / 1e6:.1f} M q/s  ms/step {d['ms_per_step']:.4f}  kernel {r['kernel_ms']:.4f}  exec frac {r['frac']:.4f}  "
      f"pairs {r['executed_pair_fraction']:.4f}  alg {r['algorithmic_tflops']:.1f} TF ({r['algorithmic_speedup']:.2f}x)  brute {r['bruteforce']['kernel_ms']:.3f} ms = {r['bruteforce']['frac']:.3f}")
e = d["e2e"]
print(f"e2e {e['value'] / 1e6:.1f} M q/s  {e['ms_per_step']:.3f} ms  d2h floor {e.get('d2h_floor_ms', 0):.3f}  ratio {e.get('d2h_floor_over_e2e', 0):.3f}  int32 {e.get('int32_idx', {}).get('ms_per_step', 0):.3f} ms")
s = d["secondary"]
print(f"chamfer {s['ms_per_step']:.4f} ms ({s['value']:.0f} pairs/s)  fresh lengths {s.get('ms_per_step_fresh_lengths', 0):.4f}  launches {s.get('launches_per_step')}")
if "sharded" in s:
    print("  sharded", {k: s['sharded'][k] for k in ('ms_per_step', 'collective_us', 'added_us_vs_local_step')})
if "secondary_ragged" in d:
    g = d["secondary_ragged"]
    print(f"ragged {g['value'] / 1e6:.1f} M q/s  {g['ms_per_step']:.4f} ms  kernel {g['kernel_ms']:.4f}")
if "secondary_highdim" in d:
    h = d["secondary_highdim"]
    print(f"highdim {h['value'] / 1e6:.2f} M q/s  {h['ms_per_step']:.3f} ms  {h['kernels_ms']}  frac {h['roofline']['frac']:.3f}")
for k in ("secondary_gather", "secondary_knn_backward", "secondary_pack"):
    if k in d:
        q = d[k]["roofline"]
        print(f"{k}: {q['kernel_ms'] * 1e3:.1f} us  frac {q['frac']:.3f}  frac_dram {q['frac_dram']:.3f}  traffic {q['traffic']}")
        if "u16" in d[k]:
            q = d[k]["u16"]
            print(f"   u16: {q['kernel_ms'] * 1e3:.1f} us  frac_dram {q['frac_dram']:.3f}")
        if "all" in d[k]:
            for a, b in d[k]["all"].items():
                for c, q in b.items():
                    print(f"   {a} {c}: {q['kernel_ms'] * 1e3:.1f} us  frac_dram {q['frac_dram']:.3f}")
if "secondary_fps" in d:
    print(f"fps {d['secondary_fps']['value']:.3f} us/iter  {d['secondary_fps']['ms_per_step']:.3f} ms")
if "secondary_ball_query" in d:
    b = d["secondary_ball_query"]
    print(f"ball query {b['ms_per_step']:.3f} ms  kernels {b['kernel_ms']:.3f}")
v = d.get("vs_reference_cuda", {})
print("vs reference CUDA:", {k: round(x["speedup"], 1) for k, x in v.items() if isinstance(x, dict) and "speedup" in x} or v.get("unavailable"))
print("clocks", d.get("clocks"), " gpu_launches", d.get("gpu_launches"), " cpu_baseline", d.get("cpu_baseline", {}).get("value"))
