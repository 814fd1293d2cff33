import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: print(l.rstrip()); continue
    print(d["case"], d["l2"], d["tuned"], "ctas", d["ctas"], "kloop", d["kloop"], "epi", d["epilogue"], d["epilogue_max"], "total", d["cta_total_max"])
