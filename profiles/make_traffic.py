#!/usr/bin/env python
"""profiles/traffic.json from the ncu summaries of one launch of each fused kernel at the bench size.

    python profiles/make_traffic.py TAG [WORLDS]     # reads profiles/TAG_ncu_{prepare,gs,finish}_WORLDS.txt

DRAM bytes = dram__bytes_read.sum + dram__bytes_write.sum; executed fp64 flops =
2 x dfma + dmul + dadd (smsp__sass_thread_inst_executed_op_*_pred_on.sum), both per world.
"""
import json
import os
import re
import sys

UNIT = {"byte": 1., "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def read(path):
    out = {"kernel": None}
    for ln in open(path):
        if ln.startswith("kernel:"):
            out["kernel"] = re.match(r"kernel:\s+(?:void\s+)?(\w+)", ln).group(1)
            continue
        t = ln.split()
        if len(t) >= 2:
            try:
                v = float(t[1])
            except ValueError:
                continue
            out[t[0]] = v*UNIT.get(t[2], 1.) if len(t) > 2 else v
    return out


def main():
    tag = sys.argv[1]
    W = int(sys.argv[2]) if len(sys.argv) > 2 else 262144
    here = os.path.dirname(os.path.abspath(__file__))
    by, fl = {}, {}
    for st in ("prepare", "gs", "finish"):
        r = read(os.path.join(here, "%s_ncu_%s_%d.txt" % (tag, st, W)))
        by[r["kernel"]] = (r["dram__bytes_read.sum"] + r["dram__bytes_write.sum"])/W
        fl[r["kernel"]] = (2*r["smsp__sass_thread_inst_executed_op_dfma_pred_on.sum"]
                           + r["smsp__sass_thread_inst_executed_op_dmul_pred_on.sum"]
                           + r["smsp__sass_thread_inst_executed_op_dadd_pred_on.sum"])/W
    doc = {"human36_contact": {
        "dram_bytes_per_world_step": sum(by.values()),
        "per_kernel_bytes_per_world": by,
        "executed_fp64_flop_per_world_step": sum(fl.values()),
        "per_kernel_executed_fp64_flop_per_world": fl,
        "source": "ncu --set full at the BENCH SIZE (%d distinct worlds, staggered-episode mix, worlds sorted by "
                  "contact state): dram__bytes_read.sum + dram__bytes_write.sum and "
                  "smsp__sass_thread_inst_executed_op_{2 x dfma, dmul, dadd}_pred_on.sum of one launch of each "
                  "fused kernel (profiles/%s_ncu_{prepare,gs,finish}_%d.txt, profiles/make_traffic.py); the sort "
                  "and the state gather/scatter (every 2nd step) add about 3 KB per world-step" % (W, tag, W)}}
    json.dump(doc, open(os.path.join(here, "traffic.json"), "w"), indent=1)
    print(json.dumps(doc, indent=1))


if __name__ == "__main__":
    main()
