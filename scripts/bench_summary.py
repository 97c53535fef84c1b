import json, sys
d = json.loads(sys.stdin.read())
print(sys.argv[1] if len(sys.argv) > 1 else "", "pairs/s", d["value"], "e2e", d["e2e"]["value"], "ms/step", d["ms_per_step"],
      {k: (v["ms"], v["tflops"]) for k, v in d["kernel_classes"].items()})
