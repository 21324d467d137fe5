import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], d["ms_per_step"], {k:round(v,3) for k,v in list(d["kernel_ms_per_step"].items())[:3]})
