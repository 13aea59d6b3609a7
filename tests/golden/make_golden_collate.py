"""Golden fixture for the batch schema: the UNMODIFIED reference's PaddingCollate
(pepflow/utils/data.py:19-78) on two synthetic items of different length that also carry the list / string fields of
a real PepDataset item (chain_id, icode, resseq, id), with and without rounding the length up to a multiple of 8.
Build-container only:  python tests/golden/make_golden_collate.py  ->  tests/golden/collate.npz"""
import json
import os
import sys
import warnings

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
warnings.filterwarnings("ignore")

import ref_shim  # noqa: E402
from make_golden import save  # noqa: E402
from pepflowww_b200.pep_dataloader import make_synthetic_complex  # noqa: E402


def items():
    out = []
    for idx, (lr, lp) in enumerate(((9, 5), (7, 4))):
        d = make_synthetic_complex(idx, lr, lp, seed=3)
        L = lr + lp
        d["chain_id"] = ["B"] * lr + ["A"] * lp
        d["icode"] = [" "] * L
        d["resseq"] = d["res_nb"].clone() + 100
        d["only_in_first"] = torch.zeros(L) if idx == 0 else None
        if d["only_in_first"] is None:
            del d["only_in_first"]
        out.append(d)
    return out


def main():
    ref_shim.load_reference()
    from pepflow.utils.data import PaddingCollate
    arrs = {}
    for tag, eight in (("e8", True), ("e1", False)):
        b = PaddingCollate(eight=eight)(items())
        for k, v in b.items():
            if isinstance(v, torch.Tensor):
                arrs[f"{tag}_{k}"] = v
            else:
                arrs[f"{tag}_{k}__json"] = torch.frombuffer(bytearray(json.dumps(v).encode()), dtype=torch.uint8)
    save("collate", **arrs)


if __name__ == "__main__":
    main()
