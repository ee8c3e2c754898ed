import json,sys
for f in sys.argv[1:]:
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f,"ERR",e); continue
    r=d.get("roofline") or {}; iso=r.get("isolated") or {}
    print(f"{f}: {d['config']['workload']} value={d['value']/1e6:.2f}M chunks/s ms/step={d['ms_per_step']*1e3:.1f}us frac={r.get('frac', 0):.3f} | isolated {iso.get('kernel_ms_mean',0)*1e3:.1f}us frac={iso.get('frac',0):.3f}")
    if d.get("e2e"): print("   e2e %.0f chunks/s"%d["e2e"]["value"], " cpu", (d.get("cpu_baseline") or {}).get("value"))
    for k,v in (d.get("also") or {}).items(): print("   also",k,{a:(round(b,4) if isinstance(b,float) else b) for a,b in v.items() if a not in ("note","api")})
    print("   clocks",d.get("clocks"))
