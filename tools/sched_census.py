#!/usr/bin/env python
"""Summarise a GGML_SCHED_DEBUG=2 log (ggml-backend.cpp:843-881, ggml_backend_sched_print_assignments): for the LAST graph the scheduler printed, how many
splits each backend got and how many nodes of which op ran where.  The evidence for 'which nodes stayed on the CPU' (SURVEY.md 8f rank 2).

  python tools/sched_census.py gpurun_out/omni_apm_sched.txt"""
import collections
import re
import sys


def census(text: str) -> dict:
    graphs = text.split("## SPLIT #0:")
    if len(graphs) < 2:
        return {"error": "no scheduler dump in the log"}
    last = "## SPLIT #0:" + graphs[-1]
    splits = collections.Counter(re.findall(r"## SPLIT #\d+: (\S+) #", last))
    nodes = collections.defaultdict(collections.Counter)
    for op, backend in re.findall(r"node #\s*\d+ \(\s*([A-Z_0-9]+)\):.{0,40}?\[\s*(\S+)\s+[^\]]*\] use=", last):
        nodes[backend][op] += 1
    return {"graphs_printed": len(graphs) - 1, "splits": dict(splits), "nodes": {b: dict(c.most_common()) for b, c in nodes.items()}}


if __name__ == "__main__":
    r = census(open(sys.argv[1], errors="replace").read())
    if "error" in r:
        sys.exit(r["error"])
    print(f"graphs printed: {r['graphs_printed']}; last graph: splits per backend {r['splits']}")
    for b, ops in r["nodes"].items():
        print(f"  {b}: {sum(ops.values())} nodes: " + ", ".join(f"{k} x{v}" for k, v in ops.items()))
