#!/usr/bin/env python
"""Where does one CLI invocation spend its time?  Times every golden command through bin/alignTools."""
import json, os, subprocess, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
g = json.load(open(os.path.join(ROOT, "tests/golden/cli_vectors.json")))
d = tempfile.mkdtemp()
for fname, info in g["files"].items():
    with open(os.path.join(d, fname), "w") as f:
        for r in info["records"]:
            f.write(">" + r["name"] + (" " + r["comment"] if r["comment"] else "") + "\n" + r["seq"] + "\n")
for v in g["vectors"]:
    if v["rc"]:
        continue
    args = [a.replace("$T", d) for a in v["argv"]]
    for env in ({}, {"CUDA_VISIBLE_DEVICES": "0"}):
        t0 = time.perf_counter()
        pr = subprocess.run([os.path.join(ROOT, "bin/alignTools")] + args, capture_output=True, env={**os.environ, **env})
        print(v["id"], " ".join(v["argv"]), env, f"{time.perf_counter() - t0:.2f}s rc={pr.returncode}", flush=True)
