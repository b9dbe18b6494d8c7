import json,sys
d=json.load(open(sys.argv[1]))
print({k:d[k] for k in ["value","n_gpus","ms_per_step","gpu_launches","clocks"]}, d["e2e"]["value"], d["roofline"]["frac"], d["phases_ms"])
