"""Import the UNMODIFIED reference (ethz-asl/multipoint) read-only from /root/reference.

Test-infrastructure only (golden-vector generation in the build container).  The reference
does not import on Python 3.12 / numpy 2.x: ``collections.Mapping`` (multipoint/utils/utils.py:22),
``np.int`` default argument (multipoint/utils/homographies.py:331) and ``np.float``
(multipoint/utils/evaluation.py:291-292) are gone.  Three aliases set *before* the import fix that
without touching the tree.  /root/reference does not exist on the GPU box: nothing that runs there
may import this module.
"""
import collections
import collections.abc
import os
import sys

import numpy as np

REFERENCE_ROOT = os.environ.get("MULTIPOINT_REFERENCE", "/root/reference")


def import_reference():
    if not os.path.isdir(os.path.join(REFERENCE_ROOT, "multipoint")):
        raise RuntimeError("reference checkout not found at %s" % REFERENCE_ROOT)
    collections.Mapping = collections.abc.Mapping
    if not hasattr(np, "int"):
        np.int = int
    if not hasattr(np, "float"):
        np.float = float
    sys.dont_write_bytecode = True  # the tree is read-only
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import multipoint.models as models
    import multipoint.utils as utils
    return models, utils
