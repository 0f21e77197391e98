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
  * parity is PINNED (tests/test_oracle_kats.py lists each source) by
      - the reference's own golden vectors (anchors, box coder, IoU, matcher, assigner, losses, NMS, PS-ROI);
      - the importable NumPy fragments of the reference (utils/np_box_ops.py, np_box_list_ops.py) run in this
        container (tests/golden/make_golden.py);
      - the reference's TF graph code EXECUTED on a NumPy stand-in for TensorFlow (tests/golden/tf_numpy_shim.py):
        core/losses.py, core/target_assigner.py (incl. the fork's closeness targets), and the meta-architecture's
        methods for the RPN post-processing / proposal sampling / all eight losses / the refiner's input assembly /
        the second-stage postprocess / position-sensitive crops (make_loss_golden.py, make_assign_golden.py,
        make_stage2_loss_golden.py, make_graph_golden.py);
      - the reference's record writers run under recording stubs for the auxiliary labels (make_aux_golden.py,
        make_coco_aux_golden.py) and its evaluator for the metrics (make_eval_golden.py);
  * what stays UNPINNED are the TF C++ kernels themselves (crop_and_resize, resize_bilinear, conv / pool SAME
    arithmetic, non_max_suppression -- cross-checked against the reference's NumPy NMS and torchvision.ops.nms),
    the L2-in-total-loss sum and the clip + momentum update: restated from their published semantics (DESIGN.md
    section 3 has the table).
"""
