"""TEST INFRASTRUCTURE: CPU oracle for the PC-GNN hot path (see port.py, pcg_oracle.c, ref_harness.py).
Nothing under pc-gnn_b200/ may import this package."""
