"""helpers shared by the golden-vector tests"""
import hashlib
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.golden.make_golden import make_image  # noqa: E402,F401  (same generator the pins were made with)


def sha(b):
    return hashlib.sha256(bytes(b)).hexdigest()


def spec_id(spec):
    return "%s%dx%d%s" % (spec["kind"], spec["w"], spec["h"], "g" if spec.get("gray") else "")
