import json,sys
for l in open(sys.argv[1]):
    j = json.loads(l)
    print(j["dims"], "win", j.get("window_planes"), "diff", j.get("max_rel_diff_vs_first"))
    for k in j:
        if k.startswith("tma="): print("   ", k, j[k]["ms_step"], j[k]["passes"])
