"""16gt_b200 -- B200-native (sm_100a) implementation of 16GT's calling hot path.

The product is native: 16gt_b200/lib/libgt16.so (CUDA kernels + C ABI, include/gt16.h) and the
drop-in 16gt_b200/bin/snapshotSnpcaller.  This package only holds the build driver and a ctypes
binding used by tests, bench.py and __graft_entry__.smoke().
"""
