"""arcticinference_b200 - B200-native (sm_100a) implementation of ArcticInference's data-parallel
hot path: the Arctic LSTM speculator draft head and SwiftKV's bulk reshape-and-cache KV write.

    from arcticinference_b200 import py_custom_ops          # torch op arctic_inference::reshape_and_cache_flash_bulk
    from arcticinference_b200.speculator import ArcticLSTMSpeculator, ArcticProposer
    from arcticinference_b200.mlp_speculator import ArcticMLPSpeculator      # MLP / sum_rnn draft head
    from arcticinference_b200.sum_rnn_speculator import ArcticSumRNNSpeculator   # sum_rnn with stacked (list) dims
    from arcticinference_b200 import swiftkv                # swiftkv_select: bulk KV write + one-launch compaction

Numeric work lives in libarctic_b200.so (C ABI: include/arctic_b200.h); build it in-tree with
``python -m arcticinference_b200._build``.  There is no CPU / PyTorch fallback.
"""
__version__ = "0.1.0"
