#!/usr/bin/env python
"""Static evidence for the kernels of the bench line: `cuobjdump -res-usage` and a SASS opcode histogram of the
built library (no GPU needed).  Writes profiles/<round>_sass_summary.json.

    python benchmarks/sass_summary.py r02
"""
import collections
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "adtomo.jl_b200", "libadtomo_b200.so")
KERNELS = {   # mangled-name fragments -> label
    "k_fwd3d_v3ILi512ELi2ELi72ELb0ELb0": "k_fwd3d_v3<512,2,72,false,false> (bench C3 forward: plain sweep, no ragged edge)",
    "k_fwd3d_v3ILi512ELi2ELi104ELb0ELb0": "k_fwd3d_v3<512,2,104,false,false> (bench C4 forward)",
    "k_fwd3d_v3ILi512ELi2ELi72ELb0ELb1": "k_fwd3d_v3<512,2,72,false,true> (ragged grids: lane mask)",
    "k_fwd3d_v3ILi512ELi2ELi72ELb1ELb0": "k_fwd3d_v3<512,2,72,true,false> (cp.async staged sweep: one CTA per SM)",
    "k_fwd3d_v4ILi512ELi2ELi72": "k_fwd3d_v4<512,2,72> (opt-in slot-block sweep)",
    "k_adj3d_sparseILi1024": "k_adj3d_sparse<1024> (bench adjoint)",
    "k_adj3d_topo2ILi1024ELb1": "k_adj3d_topo2<1024,true> (dense-rhs adjoint)",
    "k_adj3d_setup3": "k_adj3d_setup3",
    "k_fwd3d_teamILi512ELi2ELi1": "k_fwd3d_team<512,2,1>",
    "k_fwd3d_teamILi512ELi2ELi2": "k_fwd3d_team<512,2,2>",
    "k_fwd3d_teamILi512ELi1ELi1": "k_fwd3d_team<512,1,1> (at most one CTA per SM: no register cap)",
    "k_fwd3d_teamILi512ELi1ELi2": "k_fwd3d_team<512,1,2>",
    "k_adj3d_topo_teamILi512ELi4096": "k_adj3d_topo_team<512>",
    "k_model_fwd": "k_model_fwd (on-device parametrisation)",
}


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
    res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout.splitlines()
    usage = {}
    for i, ln in enumerate(res):
        m = re.search(r"Function (\S+):", ln)
        if m and i + 1 < len(res):
            usage[m.group(1)] = res[i + 1].strip()
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    out = {"library": os.path.relpath(LIB, ROOT), "command": "cuobjdump -res-usage / -sass (sm_100a cubin of the in-tree build)",
           "kernels": {}}
    blocks = re.split(r"\n\s*Function : ", sass)
    for blk in blocks[1:]:
        name = blk.split("\n", 1)[0].strip()
        label = next((v for k, v in KERNELS.items() if k in name), None)
        if not label:
            continue
        ops = collections.Counter()
        for ln in blk.splitlines():
            m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", ln)
            if m:
                ops[m.group(1)] += 1
        tot = sum(ops.values())
        fam = collections.Counter()
        for op, c in ops.items():
            if op in ("DADD", "DMUL", "DFMA", "DSETP", "DMNMX"): fam["fp64"] += c
            elif op.startswith("MUFU"): fam["mufu"] += c
            elif op in ("LDG", "STG", "LD", "ST", "ATOMG", "RED", "ATOM"): fam["global memory"] += c
            elif op in ("LDS", "STS", "ATOMS", "LDSM"): fam["shared memory"] += c
            elif op in ("LDL", "STL"): fam["local (spill)"] += c
            elif op in ("BAR", "MEMBAR", "CCTL", "ERRBAR", "DEPBAR", "WARPSYNC", "BSYNC", "BSSY"): fam["sync"] += c
            elif op in ("BRA", "EXIT", "CALL", "RET", "BRX", "JMP"): fam["branch"] += c
            elif op in ("SHFL", "VOTE", "VOTEU", "MATCH", "REDUX"): fam["warp"] += c
            elif op.startswith(("UTMA", "UBLK", "LDGSTS", "LDGDEPBAR")): fam["async copy"] += c
            else: fam["integer / move / select"] += c
        out["kernels"][label] = {"mangled": name, "res_usage": usage.get(name), "static_instructions": tot,
                                 "families": dict(fam.most_common()), "top_opcodes": dict(ops.most_common(24))}
    path = os.path.join(ROOT, "profiles", f"{tag}_sass_summary.json")
    json.dump(out, open(path, "w"), indent=1)
    print(path, {k: (v["static_instructions"], v["res_usage"]) for k, v in out["kernels"].items()})


if __name__ == "__main__":
    main()
