import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
k = d.pop('kernels', {})
print(f"ms/step {d['ms_per_step']:.4f}  value {d['value']:.3e}  e2e {d['e2e']}  step-roofline {d['roofline']['step']['frac']:.3f}  clocks {d['clocks']}")
print(f"roofline: {d['roofline']['kernel']} {d['roofline']['achieved']} GB/s frac {d['roofline']['frac']}")
print("cpu:", d.get('cpu_baseline'))
for n, v in sorted(k.items(), key=lambda kv: -kv[1]['ms_per_step']):
    print(f"  {n:28s} x{v['launches_per_step']:.0f} {v['us_per_launch']:8.1f} us {v['share']*100:5.1f}% {v.get('algorithmic_GBps', 0):8.0f} GB/s")
