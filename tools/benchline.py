"""prints the essentials of a bench.py JSON line: python tools/benchline.py gpurun_out/x.json"""
import json, sys
for f in sys.argv[1:]:
    l = json.loads(open(f).read().strip().splitlines()[-1])
    r = l.get("roofline") or {}
    print(f"{f}: {l['value']:.3f} img/s ({l['ms_per_step']:.2f} ms/image) e2e {l['e2e']['value']:.3f} launches/step {l['gpu_launches'] / l['steps']:.0f} "
          f"gemm frac {r.get('frac', 0):.3f} clocks {l.get('clocks')}")
    for k, v in (l.get("kernel_classes") or {}).items():
        if v["launches"]:
            print(f"   {k:14s} {v['launches']:6d} launches {v['ms']:8.2f} ms")
    if l.get("c5"):
        print("   c5:", l["c5"]["value"], "img/s")
