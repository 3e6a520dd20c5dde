"""sparseeventid_b200: B200-native (sm_100a) hot path of coreyjadams/SparseEventID.

Only what the submanifold sparse-convolution ResNet path needs:
  csrc/          hand-written CUDA kernels + the C ABI (include/scn_b200.h) -> lib/libscn_b200.so
  scn/           host-side mirror of the ``sparseconvnet`` module surface the reference uses
  networks.py    mirror of the reference encoder/heads (src/networks/*.py) for runs without the reference
  trainer.py     event-sharded data-parallel training step (NCCL flat-arena allreduce)
  synthetic.py / data_transforms.py   synthetic larcv-shaped events and the larcv -> SCN tuple transform
"""
__version__ = "0.1.0"
