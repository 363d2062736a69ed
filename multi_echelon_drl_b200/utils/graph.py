"""Static schedule of the information flow (reference utils/graph.py:5-44), computed once on the
host.  The kernels have the serial chain's schedule compiled in; this module derives the order from a
network description, so that a YAML/dict config can be checked against what the kernels implement."""
from __future__ import annotations

from typing import Dict, List

import yaml


def parse_order_sequence(nodes: List[str], customers_dict: Dict[str, List[str]]) -> List[str]:
    """Order in which nodes place orders: a node orders only after all of its customers did.
    Post-order depth-first walk over customers, roots visited in sorted-name order (the reference's
    deterministic tie-break), every node exactly once."""
    done, sequence = set(), []

    def visit(name: str):
        stack = [(name, iter(customers_dict.get(name, ())))]
        while stack:
            cur, it = stack[-1]
            nxt = next((c for c in it if c not in done and all(c != s[0] for s in stack)), None)
            if nxt is not None:
                stack.append((nxt, iter(customers_dict.get(nxt, ()))))
            else:
                stack.pop()
                if cur not in done:
                    done.add(cur)
                    sequence.append(cur)

    for root in sorted(nodes):
        if root not in done:
            visit(root)
    return sequence


def read_yaml(file_path: str) -> dict:
    with open(file_path, "r") as fh:
        return yaml.safe_load(fh)


def chain_from_config(network_config: dict):
    """Validate that a network description (network_configs/*.yaml layout: nodes with holding_cost /
    backorder_cost / demand_path / is_external_supplier, arcs with customer / supplier / info_leadtime /
    shipment_leadtime) is the serial chain the kernels implement and return
    (node names downstream->upstream, info lead times, shipment lead times, holding, backorder).
    A node with a `demand_path` is the demand source (the reference's loader inverts this test and
    crashes, supplynetwork.py:360-365, Q1)."""
    nodes = {n["name"]: n for n in network_config["nodes"]}
    arcs = network_config["arcs"]
    customers = {name: [a["customer"] for a in arcs if a["supplier"] == name] for name in nodes}
    suppliers = {name: [a["supplier"] for a in arcs if a["customer"] == name] for name in nodes}
    internal = [n for n in nodes if not nodes[n].get("is_external_supplier", False)]
    seq = [n for n in parse_order_sequence(list(nodes), customers) if n in internal]
    if any(len(customers[n]) > 1 or len(suppliers[n]) != 1 for n in internal):
        raise NotImplementedError("only serial chains are supported (one supplier and at most one customer per node)")
    sources = [n for n in internal if nodes[n].get("demand_path")]
    if sources != seq[:1]:
        raise NotImplementedError("the demand source must be the most downstream node of the chain")
    for a, b in zip(seq, seq[1:]):
        if suppliers[a] != [b]:
            raise NotImplementedError("nodes do not form a single chain")
    if not nodes[suppliers[seq[-1]][0]].get("is_external_supplier", False):
        raise NotImplementedError("the most upstream node must buy from an external supplier")
    arc_of = {a["customer"]: a for a in arcs}
    info = tuple(int(arc_of[n].get("info_leadtime", 0)) for n in seq)
    ship = tuple(int(arc_of[n].get("shipment_leadtime", 0)) for n in seq)

    def per_item(v):
        return list(v) if isinstance(v, (list, tuple)) else [v]

    n_items = max(1, len(network_config.get("items", []) or []))
    hold = [[per_item(nodes[n].get("holding_cost", 0))[min(i, len(per_item(nodes[n].get("holding_cost", 0))) - 1)]
             for n in seq] for i in range(n_items)]
    back = [[per_item(nodes[n].get("backorder_cost", 0))[min(i, len(per_item(nodes[n].get("backorder_cost", 0))) - 1)]
             for n in seq] for i in range(n_items)]
    return seq, info, ship, hold, back
