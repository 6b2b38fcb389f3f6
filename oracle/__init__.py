"""TEST INFRASTRUCTURE ONLY - never imported by the product path.

``oracle/`` holds CPU restatements of the reference's algorithm for the two
halves of the hot path:

* ``oracle.speculator_ref``  - Arctic LSTM speculator draft-head forward
  (reference: arctic_inference/vllm/spec_dec/arctic_speculator.py:59-95, 648-866)
* ``oracle.kv_write_ref``    - SwiftKV bulk reshape-and-cache
  (reference: csrc/custom_ops/kernels.cu:12-69, 86-155), numpy + plain C

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package, and only as the checker or the
timed CPU baseline.  ``arcticinference_b200`` must never import it.
"""
