"""Drop-in ``sparseconvnet`` package: ``import sparseconvnet as scn`` in the unmodified reference
(src/networks/resnet.py:2, src/networks/sparse_building_blocks.py:3, src/networks/torch/sparseresnet*.py,
bin/sparse_efficiency.py:6) resolves to the B200-native implementation."""
from sparseeventid_b200.scn import *          # noqa: F401,F403
from sparseeventid_b200.scn import __all__    # noqa: F401
