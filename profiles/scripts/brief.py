import json,sys
d=json.load(open(sys.argv[1])); print(round(d["value"]), round(d["e2e"]["value"]), round(d["us_per_update"],1))
