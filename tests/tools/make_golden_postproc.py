"""Generate tests/golden/postproc.npz from the UNMODIFIED reference FCOSPostProcessor (build container only).

    PYTHONPATH=. python tests/tools/make_golden_postproc.py

The reference's NMS is the compiled `fcos_core._C.nms` (csrc/cuda/nms.cu), which cannot be built here (THC headers); the shim's
`_C` stub gets `nms` = oracle.condgraph_oracle.nms_indices (a restatement of nms.cu's rule) BEFORE fcos_core.layers is imported.
Everything else -- candidate selection, top-k, decoding, clipping, per-class loop, kthvalue cap -- is the reference's own code
(inference.py:54-194, run in TEST.MODE 'precision' form: box_cls are probabilities).  So this fixture pins the post-processor
EXCEPT the NMS rule itself, which stays pinned only by the citation.
"""
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import condgraph_oracle as orc  # noqa: E402
from oracle import ref_shim  # noqa: E402
import postproc_case  # noqa: E402


def main():
    ref_shim.install()
    sys.modules["fcos_core._C"].nms = lambda boxes, scores, thr: orc.nms_indices(boxes, scores, thr)
    from fcos_core.modeling.rpn.fcos.inference import FCOSPostProcessor
    out = {}
    for name in postproc_case.CASES:
        d = postproc_case.build(name)
        pp = FCOSPostProcessor(pre_nms_thresh=d["thr"], pre_nms_top_n=d["top_n"], nms_thresh=d["nms_thr"], fpn_post_nms_top_n=d["post_n"],
                               min_size=0, num_classes=d["num_fg"] + 1, mode="precision")
        locations = [torch.stack(orc.level_locations(h, w, s), dim=1) for (h, w), s in zip(d["shapes"], d["strides"])]
        # torch-version drift: inference.py:76 calls .view(N, -1) on `box_cls > thresh`, which torch 2.x leaves with the permuted
        # strides when C == 1; handing the maps over in channels-last memory makes the reference's own permute contiguous
        # (same values, same shapes)
        cl = [p.contiguous(memory_format=torch.channels_last) for p in d["probs"]]
        res = pp(locations, cl, d["regs"], d["ctrs"], d["sizes"])
        for i, bl in enumerate(res):
            c = postproc_case.canonical(bl.bbox.numpy(), bl.get_field("scores").numpy(), bl.get_field("labels").numpy())
            out["%s/img%d" % (name, i)] = c
            # the reference's own order too (valid wherever no level exceeded pre_nms_top_n)
            out["%s/img%d_labels_in_order" % (name, i)] = bl.get_field("labels").numpy().astype(np.int64)
            out["%s/img%d_scores_in_order" % (name, i)] = bl.get_field("scores").numpy()
        print(name, [len(bl) for bl in res])
    meta = "reference=/root/reference FCOSPostProcessor (nms = oracle restatement of nms.cu) torch=%s" % torch.__version__
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "postproc.npz"), __meta__=np.array(meta), **out)


if __name__ == "__main__":
    main()
