"""Dependency-free reader of a TF-1 ``MetaGraphDef`` (.meta): walks the protobuf wire format and
returns {node name: (op, inputs, attrs)}.  TEST INFRASTRUCTURE ONLY -- used once, in this container,
to pull the constants of the graph the authors actually ran out of the shipped checkpoint
(``python -m oracle.metagraph`` writes tests/golden/graph_constants.json)."""
from __future__ import annotations

import json
import os
import struct
import sys


def _varint(b, i):
    r = s = 0
    while True:
        c = b[i]
        i += 1
        r |= (c & 0x7F) << s
        if c < 0x80:
            return r, i
        s += 7


def fields(b):
    i, n = 0, len(b)
    while i < n:
        tag, i = _varint(b, i)
        f, w = tag >> 3, tag & 7
        if w == 0:
            v, i = _varint(b, i)
        elif w == 1:
            v, i = b[i:i + 8], i + 8
        elif w == 2:
            ln, i = _varint(b, i)
            v, i = b[i:i + ln], i + ln
        elif w == 5:
            v, i = b[i:i + 4], i + 4
        else:
            raise ValueError(f"wire type {w}")
        yield f, w, v


def _tensor(b):
    out = {"dtype": None, "shape": [], "vals": []}
    for f, w, v in fields(b):
        if f == 1:
            out["dtype"] = v
        elif f == 2:
            for f2, _, v2 in fields(v):
                if f2 == 2:
                    out["shape"].append(next((x for ff, _, x in fields(v2) if ff == 1), 0))
        elif f == 4:
            fmt = {1: "f", 2: "d", 3: "i", 9: "q"}.get(out["dtype"])
            if fmt:
                out["vals"] = list(struct.unpack(f"<{len(v) // struct.calcsize(fmt)}{fmt}", v))
        elif f == 5:
            out["vals"] += list(struct.unpack(f"<{len(v) // 4}f", v)) if w == 2 else [struct.unpack("<f", v)[0]]
        elif f == 7:
            if w == 0:
                out["vals"].append(v if v < 1 << 63 else v - (1 << 64))
            else:
                j = 0
                while j < len(v):
                    x, j = _varint(v, j)
                    out["vals"].append(x)
    return out


def _attr(b):
    for f, w, v in fields(b):
        if f == 8:
            return _tensor(v)
        if f == 3:
            return v if v < 1 << 63 else v - (1 << 64)
        if f == 4:
            return struct.unpack("<f", v)[0]
        if f == 5:
            return bool(v)
        if f == 2:
            return v.decode(errors="replace")
    return None


def read_nodes(path):
    meta = open(path, "rb").read()
    graph = next(v for f, _, v in fields(meta) if f == 2)
    nodes = {}
    for f, _, nd in fields(graph):
        if f != 1:
            continue
        name = op = None
        inputs, attrs = [], {}
        for f2, _, v in fields(nd):
            if f2 == 1:
                name = v.decode()
            elif f2 == 2:
                op = v.decode()
            elif f2 == 3:
                inputs.append(v.decode())
            elif f2 == 5:
                k = val = None
                for f3, _, v3 in fields(v):
                    if f3 == 1:
                        k = v3.decode()
                    elif f3 == 2:
                        val = _attr(v3)
                attrs[k] = val
        nodes[name] = (op, inputs, attrs)
    return nodes


def const(nodes, name):
    op, _, attrs = nodes[name]
    assert op == "Const", (name, op)
    v = attrs["value"]["vals"]
    return v[0] if len(v) == 1 else v


META = "/root/reference/ckpt_DeepMimicWalk-v0/deepmimic_dppo_pfpn_particle35/34114/model.ckpt-78000.meta"


def main():
    nodes = read_nodes(META)
    ops = {}
    for n, (op, _, _) in nodes.items():
        ops[op] = ops.get(op, 0) + 1

    def find(prefix, op):
        return sorted(n for n, (o, _, _) in nodes.items() if n.startswith(prefix) and o == op)

    def consts_under(prefix):
        return {n: const(nodes, n) for n in find(prefix, "Const")
                if isinstance(const(nodes, n), (int, float)) or len(const(nodes, n)) <= 4}

    out = {
        "source": META.replace("/root/reference/", ""),
        "n_nodes": len(nodes),
        "normal_prob_consts": consts_under("global_net/actor/Normal/prob"),
        "pi_target_consts": consts_under("global_net/policy_loss/pi_target"),
        "clipped_surrogate_consts": consts_under("global_net/policy_loss/clipped_surrogate"),
        "normalize_advantage_consts": consts_under("global_net/normalize_advantage"),
        "state_clip_consts": consts_under("global_net/state_normalizer"),
        # float constants of the resampler sub-graph (threshold, +-1e-4 noise floor, logstd clip ...)
        "resample_cond_consts": {n: const(nodes, n) for n in find("cond/", "Const")
                                 if nodes[n][2]["value"]["dtype"] == 1 and isinstance(const(nodes, n), float)},
        "resample_interval": [const(nodes, n) for n in nodes if n.endswith("GreaterEqual/y")],
        "multinomial": {n: {"num_samples_input": nodes[n][1], "seed": nodes[n][2].get("seed"),
                            "seed2": nodes[n][2].get("seed2")} for n in nodes if nodes[n][0] == "Multinomial"},
        "softmax_nodes": [n for n in nodes if nodes[n][0] == "Softmax" and n.startswith("global_net")],
        "adam": {n: const(nodes, n) for n in nodes if nodes[n][0] == "Const" and
                 any(n.endswith(s) for s in ("optimizer/lr", "optimizer/beta1", "optimizer/beta2", "optimizer/epsilon"))},
        "clip_by_global_norm_consts": consts_under("optimizer/clip_by_global_norm"),
        "op_histogram_top": dict(sorted(ops.items(), key=lambda kv: -kv[1])[:25]),
    }
    dst = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "graph_constants.json")
    json.dump(out, open(dst, "w"), indent=1, sort_keys=True, default=str)
    print("wrote", dst)


if __name__ == "__main__":
    main()
