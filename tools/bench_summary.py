"""Print a one-line summary of a bench.py JSON line read from stdin (label = argv[1])."""
import json
import sys

label = sys.argv[1] if len(sys.argv) > 1 else ""
for line in sys.stdin:
    line = line.strip()
    if not line.startswith("{"):
        continue
    d = json.loads(line)
    r = d.get("roofline", {})
    pipe = r.get("pipeline_cycles_per_board", {})
    print(label, "sims/s", round(d["value"]), "ms/step", round(d["ms_per_step"], 1), "e2e", round(d["e2e"]["value"]),
          "clk", d.get("clocks", {}).get("sm_mhz"), "conv TF", round(r.get("achieved", 0), 1), "frac", round(r.get("frac", 0) or 0, 3),
          "ms/launch", round(r.get("avg_launch_ms", 0), 3), "trunk share", round(r.get("trunk_share_of_step", 0) or 0, 3),
          "wait_full", pipe.get("mma_wait_full"), "wait_tempty", pipe.get("mma_wait_tmem_empty"),
          "dropped", d.get("search", {}).get("dropped_trees"), "err", d.get("search", {}).get("games_in_error"),
          "epi", {k[4:]: v for k, v in pipe.items() if k.startswith("epi_")})
    for t, v in r.get("pipeline_cycles_per_board_by_layer_type", {}).items():
        print("   ", t, v)
