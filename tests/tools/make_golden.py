"""Generate tests/golden/*.npz from the UNMODIFIED reference (run in the build container only).

    PYTHONPATH=. python tests/tools/make_golden.py [case ...]

Imports /root/reference through oracle/ref_shim.py, drives it through tests/harness.py:run_case
(seeded inputs, fixture weights, dropout off) and stores the reduced results.  The fixtures are
what pins the oracle and the CUDA path to the reference on the GPU box, where /root/reference
does not exist.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import ref_shim  # noqa: E402
from scan_b200.config import to_plain  # noqa: E402
import harness  # noqa: E402


def main(argv):
    names = argv or list(harness.CASES)
    _, _, BoxList = ref_shim.reference_modules()
    torch.set_num_threads(os.cpu_count())
    outdir = os.path.join(ROOT, "tests", "golden")
    for name in names:
        cfg = harness.build_case(name)[0]
        ref = ref_shim.build_reference(to_plain(cfg))
        res = harness.run_case(name, ref, "reference", boxlist_cls=BoxList)
        meta = "reference=/root/reference torch=%s numpy=%s" % (torch.__version__, np.__version__)
        np.savez_compressed(os.path.join(outdir, name + ".npz"), __meta__=np.array(meta), **res)
        print(name, len(res), "arrays",
              {k: (float(v) if v.ndim == 0 else v.shape) for k, v in res.items() if "loss" in k})


if __name__ == "__main__":
    main(sys.argv[1:])
