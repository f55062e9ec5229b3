import json,sys
for path in sys.argv[1:]:
    try:
        d=json.loads(open(path).read().strip().splitlines()[-1])
    except Exception as e:
        print(path, "unreadable", e); continue
    print(path)
    print("  ", {k:d.get(k) for k in ("value","ms_per_step","step_tflops_per_gpu","gpu_launches")})
    print("  ", d.get("pct_bf16_tc_peak"), d.get("clocks"))
    r=d.get("roofline") or {}
    print("   roofline", {k:r.get(k) for k in ("achieved","frac","share_of_step","launches_timed")})
    print("   e2e", d.get("e2e"))
