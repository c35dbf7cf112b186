"""detection_3d_b200: B200-native (sm_100a) implementation of the SparseConvNet backbone path of
zhupan007/Detection_3D behind the reference's `sparseconvnet` Python API."""
import sys


def install_as_sparseconvnet():
    """Make `import sparseconvnet` resolve to this implementation (drop-in for Detection_3D)."""
    from . import sparseconvnet as scn
    sys.modules['sparseconvnet'] = scn
    sys.modules['sparseconvnet.SCN'] = scn.SCN
    return scn
