"""Generates the golden fixtures under tests/golden/ from the CPU oracle.

The reference holds no golden vectors for this path and cannot run offline
(SURVEY.md §8c), so these pin OUR arithmetic spec: any change to the oracle or
to the CUDA engine that alters a single swap decision, uniform, log ratio or
adapted schedule shows up as a diff against these files.

    python tests/golden/make_golden.py
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import numpy as np            # noqa: E402

import pigeons_jl_b200 as pg  # noqa: E402
from golden_cases import GOLDEN_CASES, summarise  # noqa: E402
from oracle_adapter import load_oracle            # noqa: E402


def main():
    lib = load_oracle()
    # <name>.json: the default recorder order (per replica + tree merge, the reference's reduce_recorders!);
    # <name>.per_chain.json: recorder_order = PGN_RECORDERS_PER_CHAIN
    for name, kw in GOLDEN_CASES.items():
        for order, suffix in ((0, ".json"), (1, ".per_chain.json")):
            if order == 1 and name.startswith("two_legs"):
                continue      # two legs need the per-replica recorder order
            pt = pg.pigeons(engine_lib=lib, record=[pg.index_process, pg.swap_trace], recorder_order=order, **kw())
            with open(os.path.join(HERE, name + suffix), "w") as f:
                json.dump(summarise(pt), f, indent=1)
            pt.close()
            print("wrote", name + suffix)


if __name__ == "__main__":
    main()
