"""CPU oracle: a literal restatement of the reference's algorithm for the hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under mtl_ssl_b200/ imports this package; it is used
by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs as
the checker and the reported CPU baseline, never as the product path.

The reference (wonheeML/mtl-ssl, /root/reference) is Python-2.7 + TensorFlow-1.7 graph
code and cannot run in this image (no TF, no python2, no protoc), and the arithmetic of
its hot ops lives in TensorFlow 1.7.0 (requirements.txt:26), which is not vendored.  So:
  * box/matcher/assigner/sampler/NMS/loss logic is restated from the reference's own
    Python files (each function cites file:line);
  * TF kernels (crop_and_resize, non_max_suppression, resize_bilinear, conv2d SAME
    padding, softmax CE, dynamic_stitch) are restated from TF 1.7's published semantics;
  * parity is PINNED for the primitives by the reference's own golden vectors
    (tests/test_oracle_kats.py lists each reference test file:line) and by running the
    importable NumPy fragments of the reference (utils/np_box_ops.py,
    utils/np_box_list_ops.py) in this container to generate tests/golden/*.npz
    (script: tests/golden/make_golden.py);
  * the fork's aux heads (window / closeness / edgemask / refine) have no reference
    tests: for those rows the oracle is "parity unpinned" and is only a literal
    line-by-line restatement.
"""
