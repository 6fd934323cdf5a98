"""Extracts the few theta_0 entries the BASELINE configs use from the reference's data file
/root/reference/examples/config/config.json (84 keys `<SYS>_<MAX|MIN><n>` -> [s_f, l_1..l_d], natural units)
into gpr.jl_b200/theta0_config.json (workload data of the package: data.py and bench.py read it).  Run in the build container (the GPU box has no /root/reference)."""
import json, os
KEYS = ["P1_MAX256", "P2_MAX1024", "CP_MAX2048", "CP_MAX512", "FB_MAX512", "P1_MAX64", "P2_MAX256"]
src = json.load(open("/root/reference/examples/config/config.json"))
out = {k: src[k] for k in KEYS}
dst = os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "gpr.jl_b200", "theta0_config.json")
json.dump(out, open(dst, "w"), indent=0)
print("wrote", dst, {k: len(v) for k, v in out.items()})
