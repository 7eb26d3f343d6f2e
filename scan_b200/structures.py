"""Minimal stand-in for the reference's BoxList (fcos_core/structures/bounding_box.py:9).

The middle head only touches `.bbox` ([G,4] xyxy fp32), `.mode`, `.get_field("labels")`
and `.area()` (loss.py:308-311).  The real `fcos_core` BoxList is accepted as-is
(duck typing); this class exists so that tests and the benchmark do not need fcos_core.
"""
import torch


class BoxList(object):
    def __init__(self, bbox, image_size, mode="xyxy"):
        bbox = torch.as_tensor(bbox, dtype=torch.float32)
        if bbox.ndimension() != 2 or bbox.size(-1) != 4:
            raise ValueError("bbox should be [G,4], got %s" % (tuple(bbox.shape),))
        if mode not in ("xyxy", "xywh"):
            raise ValueError("mode should be 'xyxy' or 'xywh'")
        self.bbox = bbox
        self.size = image_size  # (width, height)
        self.mode = mode
        self.extra_fields = {}

    def add_field(self, field, field_data):
        self.extra_fields[field] = field_data

    def get_field(self, field):
        return self.extra_fields[field]

    def has_field(self, field):
        return field in self.extra_fields

    def fields(self):
        return list(self.extra_fields.keys())

    def area(self):
        # bounding_box.py:226-236 (the +1 convention is part of the parity contract)
        box = self.bbox
        if self.mode == "xyxy":
            return (box[:, 2] - box[:, 0] + 1) * (box[:, 3] - box[:, 1] + 1)
        return box[:, 2] * box[:, 3]

    def to(self, device):
        out = BoxList(self.bbox.to(device), self.size, self.mode)
        for k, v in self.extra_fields.items():
            out.add_field(k, v.to(device) if hasattr(v, "to") else v)
        return out

    def __len__(self):
        return self.bbox.shape[0]
