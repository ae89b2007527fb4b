"""Golden vectors for balancing coolers in SQUARE storage mode (pixels on both sides of the
diagonal, ``storage-mode = "square"``; the reference's ``toy.asymm.*`` fixtures are of this kind),
made by the UNMODIFIED reference (src/cooler/_balance.py:263-477) through oracle/refshim.py:

    python tests/golden/make_square_golden.py   ->  square_tables.npz, balance_square_golden.{npz,json}

``_marginalize`` (_balance.py:63-69) adds every stored pixel to its row bin AND its column bin
whatever the storage mode, and ``_zero_diags`` tests ``abs(bin1 - bin2)``: a product path that
assumed bin2 >= bin1 anywhere (filters, transposition, column strips, 2-D blocks) fails these cases.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

from oracle import refshim  # noqa: E402

from make_golden import jsonable, random_table, table  # noqa: E402

CASES = [("gw", {}), ("cis", {"cis_only": True}), ("trans", {"trans_only": True}),
         ("gw_diag0", {"ignore_diags": 0, "mad_max": 3}), ("cis_diag1", {"cis_only": True, "ignore_diags": 1}),
         ("gw_count30", {"min_count": 30, "min_nnz": 4})]


def main():
    assert refshim.available(), "reference not present"
    inputs, bal, meta = {}, {}, {}
    for seed, nb, dens in [(21, [60, 35, 1, 44], 0.35), (22, [150, 9, 120, 70, 2], 0.2)]:
        t = random_table(seed, nb, 1000, dens, 30, symmetric=False)
        assert (t["bin2_id"] < t["bin1_id"]).any() and (t["bin2_id"] > t["bin1_id"]).any()
        name = f"sq{seed}"
        refshim.register(name + ".cool", *table(t))
        for k, v in t.items():
            inputs[f"{name}/{k}"] = v
        for case, kw in CASES:
            bias, stats, passes, warned = refshim.ref_balance(name + ".cool", **kw)
            bal[f"{name}/{case}"] = bias
            meta[f"{name}/{case}"] = dict(kwargs=kw, stats=jsonable(stats), passes=passes, warned=warned)
            print(name, case, passes, int(np.isnan(bias).sum()), len(t["count"]))
    np.savez_compressed(os.path.join(HERE, "square_tables.npz"), **inputs)
    np.savez_compressed(os.path.join(HERE, "balance_square_golden.npz"), **bal)
    with open(os.path.join(HERE, "balance_square_golden.json"), "w") as f:
        json.dump(meta, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
