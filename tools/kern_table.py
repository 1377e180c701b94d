"""Developer tool: print the per-kernel table of a bench JSON line (python tools/kern_table.py file.json [n])."""
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
n = int(sys.argv[2]) if len(sys.argv) > 2 else 8
print("value %.2f  e2e %.2f  ms/step %.3f  clocks %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d.get("clocks", {}).get("sm_mhz")))
for k in d.get("kernels", [])[:n]:
    print("  %-28s %.4f ms x %4d  share %.3f  hbm %.3f  tensor %.3f" % (k["kernel"], k["ms_per_launch"], k["launches"], k["share"], k["hbm_frac"], k["tensor_frac_useful"]))
