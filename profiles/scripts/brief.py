import json,sys
d=json.load(open(sys.argv[1])); print(round(d["value"]), round(d["e2e"]["value"]), round(d["us_per_update"],1))
if len(sys.argv) > 2:
    b = json.load(open(sys.argv[2]))
    print("graph-only replay us:", round(b.get("graph_replay_only_us", 0), 1), " sum of launches us:", round(sum(u for _, u in b["launch_us"]), 1))
