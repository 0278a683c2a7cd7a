"""pcgnn_b200 — B200-native pick-and-choose message passing (PC-GNN hot path).

Submodules mirror the reference's module names so its callers resolve unchanged:
  layers     InterAgg1/3/5, IntraAgg, choose_step_neighs, choose_step_test   (src/layers.py)
  graphsage  MeanAggregator/Encoder/GraphSage, GCNAggregator/GCNEncoder/GCN  (src/graphsage.py)
  model      PCALayer                                                        (src/model.py)
  utils      pick_step, pos_neg_split, normalize, sparse_to_adjlist_for_train (src/utils.py)
plus graph (stacked CSR), synth (synthetic datasets), parallel (target sharding + NCCL),
and _lib (ctypes binding of the C-ABI library built from csrc/).
"""
__version__ = "0.1.0"
