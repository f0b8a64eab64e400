"""Pretty-print the sub-records of a bench.py JSON line (file argument)."""
import json, sys
l = [x for x in open(sys.argv[1]) if x.startswith("{")]
d = json.loads(l[-1])
print("value", round(d["value"], 1), "ms", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"], 1), "n_gpus", d["n_gpus"], "clocks", d.get("clocks"))
print("config", d["config"].get("allreduce"), d["config"].get("dw_dtype"))
print("check", d.get("check"))
r = d["roofline"]
print("roof", {k: r[k] for k in ("bound", "frac", "kernel", "per_op_ms", "per_op_ms_warm_l2")})
for k, v in d.get("density_sweep", {}).items():
    print(k, {o: (round(v[o]["ms"], 4), round(v[o]["tflops"]), round(v[o]["frac_tensor_peak"], 3), round(v[o]["frac_hbm_peak"], 2)) for o in ("fprop", "bprop", "updat")})
for k, v in d.get("variants", {}).items():
    print(k, {o: (round(v[o]["ms"], 4), round(v[o]["tflops"]), v[o]["kernel"]) for o in ("fprop", "bprop", "updat")})
for k, v in d.get("cfg4_block_size_sweep", {}).get("results", {}).items():
    print(k, {o: (round(v[o]["ms"], 4), round(v[o]["tflops"], 1), v[o]["kernel"]) for o in ("fprop", "bprop", "updat")})
c3 = d.get("cfg3_attention")
if c3:
    print({k: (round(v["ms"], 4), round(v["frac_hbm_peak"], 2), v["kernel"]) for k, v in c3.items() if isinstance(v, dict)}, "fwd chain", c3["forward_chain_ms"])
print("cfg5", d.get("cfg5_strong"))
