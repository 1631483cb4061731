import json,sys
for line in open(sys.argv[1]):
    line=line.strip()
    if not line.startswith('{'): continue
    d=json.loads(line)
    r=d.get("roofline") or {}
    e=d.get("e2e") or {}
    print("value %.4g rows/s  ms/step %.4f  launches %s | kernel ms %s frac %s | two_call %s | e2e ms %s value %s | clocks %s" % (d["value"], d["ms_per_step"], d.get("gpu_launches"), r.get("avg_launch_ms"), r.get("frac"), (r.get("two_call_unfused") or {}).get("ms_per_step"), e.get("ms_per_step"), e.get("value"), d.get("clocks")))
