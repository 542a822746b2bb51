import json,sys
for l in open(sys.argv[1]):
    try: r = json.loads(l)
    except Exception: continue
    if "B" in r: print({k: (round(r[k],1) if isinstance(r[k], float) else r[k]) for k in ("K","B","team","ms","fits_per_s")})
    else:
        bad = (r.get("dchi2") or 0) > 1e-10 or r.get("conv", 1) < 0.99
        print("BAD" if bad else "ok ", r["K"], r["ny"], r["kind"], r["team"], r.get("polish"), r.get("dp_sd"))
