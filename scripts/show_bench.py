"""Compact view of bench.py JSON lines: python scripts/show_bench.py file.json [...]"""
import json
import sys

for f in sys.argv[1:]:
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:  # noqa: BLE001
        print(f, "unreadable:", e)
        continue
    r, e2e, cpu = d.get("roofline") or {}, d.get("e2e") or {}, d.get("cpu_baseline") or {}
    print(f"{f}: {d.get('value', 0):.4g} {d.get('unit')}  ms/step {d.get('ms_per_step', 0):.2f}  kernel ms {r.get('kernel_ms_per_launch', 0):.2f}"
          f"  frac {r.get('frac', 0):.4f}  alg B/sim {r.get('algorithmic_bytes_per_sim', 0):.0f}  e2e {e2e.get('value', 0):.4g}"
          f"  cpu {cpu.get('value', 0):.4g} ({cpu.get('cores')} cores)  B/game {d.get('bytes_per_game')}  clocks {d.get('clocks')}")
