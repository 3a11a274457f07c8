"""oracle/ -- TEST INFRASTRUCTURE, not product code.

CPU restatements of the reference's hot path, used only as the *checker*:
  * ``index_oracle``  -- ctypes doors onto oracle/libelo_oracle.so (this repo's C restatement of the
    two neighbour-search kernels) and onto oracle/_ref/libref_{cpu,gpu}.so (the reference's own
    kernels compiled from /root/reference by oracle/Makefile).
  * ``graph_oracle``  -- torch-CPU (fp32 / fp64) restatement of the TensorFlow graph blocks
    (set-conv, set-upconv, cost volume, re-projection, embedding mask, pose heads, full forward).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this package.  Nothing under efficientlo-net_b200/ does.
"""
