"""Markdown tables of the round-2 measurements from the JSON lines committed under profiles/ (python tools/r02_table.py)."""
import glob
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = os.path.join(ROOT, "profiles")


def load(name):
    try:
        return json.load(open(os.path.join(P, name)))
    except Exception:
        return None


def k(v):
    return "—" if v is None else (f"{v / 1e6:.3f} M" if v >= 1e6 else f"{v / 1e3:.1f} k")


rows = [
    ("lidar_only (configs[1]), float32 q8 in; full line of the closing one-GPU session", [("r02_final_bench.json", 1)]),
    ("lidar_only (configs[1]), N = 1 / 8 on ONE 8-GPU box, closing kernels", [("r02_lidar_1gpu_closing.json", 1), ("r02_lidar_8gpu_closing.json", 8)]),
    ("lidar_only (configs[1]), N = 1 / 2 / 4 / 8 on ONE 8-GPU box (before the front's own tile width: + 2.5 % since)", [("r02_lidar_1gpu.json", 1), ("r02_lidar_2gpu.json", 2), ("r02_lidar_4gpu.json", 4), ("r02_lidar_8gpu.json", 8)]),
    ("lidar_only, uint16 in (device resident)", [("r02_final_bench_u16_input.json", 1)]),
    ("lidar_only, arbitrary float in (dictionary path)", [("r02_final_bench_float_rank.json", 1), ("r02_float_8gpu.json", 8)]),
    ("guided (configs[2])", [("r02_final_bench_guided.json", 1), ("r02_guided_8gpu.json", 8)]),
    ("guided, arbitrary float in", [("r02_final_bench_guided_float_rank.json", 1)]),
    ("DC_lidar_camera chain (configs[2]: SLIC on the Lab image -> guided completion)", [("r02_final_bench_lidar_camera_chain.json", 1)]),
    ("stereo refinement (configs[3], a4-a9)", [("r02_final_bench_stereo.json", 1), ("r02_stereo_8gpu.json", 8)]),
    ("stereo chain (configs[3]: projection -> guided float -> refinement), closing one-GPU session", [("r02_final_bench_stereo_chain.json", 1)]),
    ("stereo chain, N = 1 / 8 on ONE 8-GPU box", [("r02_chain_1gpu_same_box.json", 1), ("r02_chain_8gpu.json", 8)]),
    ("sweep 352x1216 @ 1 %", [("r02_sweep_352x1216_p01_8gpu.json", 8)]),
    ("sweep 2048x4096 @ 1 %", [("r02_sweep_2048x4096_p01_8gpu.json", 8)]),
    ("sweep 2048x4096 @ 20 %", [("r02_sweep_2048x4096_p20_8gpu.json", 8)]),
]
print("| workload | N | frames/s (device resident) | per GPU | efficiency vs N=1 | frac of HBM roofline | e2e frames/s | copy ceiling | reference CPU frames/s |")
print("|---|---|---|---|---|---|---|---|---|")
for name, files in rows:
    base = None
    for f, n in files:
        d = load(f)
        if not d:
            continue
        v = d["value"]
        if n == 1:
            base = v
        e = d.get("e2e") or {}
        cpu = (d.get("cpu_baseline") or {}).get("value")
        eff = f"{v / n / base:.3f}" if base and n > 1 else ("1" if n == 1 else "—")
        print(f"| {name} | {n} | {k(v)} | {k(v / n)} | {eff} | {d['roofline']['frac']:.4f} | {k(e.get('value'))} | {k(e.get('copy_ceiling'))} | {'—' if cpu is None else f'{cpu:.1f}'} |")
