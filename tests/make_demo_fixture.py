"""(Fixture script, not a test.)  tests/golden/demo_nag.pt: the reference's demo partition
(/root/reference/notebooks/demo_nag_v3.h5, a 4-level S3DIS room: 41 568 points, 1 192 / 501 /
166 superpoints) as read by superpoint_transformer_b200.io — build container only:
    python tests/make_demo_fixture.py
Integers are kept in the file's own (smallest) dtypes to keep the fixture at the size of the
file; tests cast with `.long()`."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from superpoint_transformer_b200.data import Cluster          # noqa: E402
from superpoint_transformer_b200.io import load_nag           # noqa: E402

SRC = '/root/reference/notebooks/demo_nag_v3.h5'
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'demo_nag.pt')


def to_plain(nag):
    levels = []
    for i in nag.level_range:
        lv = {}
        for k in nag[i].keys:
            v = nag[i][k]
            lv[k] = {'pointers': v.pointers, 'points': v.points} if isinstance(v, Cluster) else v
        levels.append(lv)
    return {'start': nag.start_i_level, 'levels': levels}


if __name__ == '__main__':
    torch.save(to_plain(load_nag(SRC)), OUT)
    print(OUT, os.path.getsize(OUT))
