"""Argument sets shared by tests/golden/make_postprocess_golden.py and the tests."""
PP_CFGS = {"multi": dict(conf_thres=0.25, iou_thres=0.45, multi_label=True, max_det=1500),
           "best": dict(conf_thres=0.3, iou_thres=0.4, multi_label=False, max_det=1000),
           "cls": dict(conf_thres=0.25, iou_thres=0.45, multi_label=True, classes=[1, 5, 9], max_det=50),
           "agn": dict(conf_thres=0.28, iou_thres=0.2, multi_label=True, agnostic=True, max_det=1500),
           "lowconf": dict(conf_thres=0.001, iou_thres=0.6, multi_label=True, max_det=300)}
PP_ANCHORS = {"lowconf": 100}
# > max_nms (30 000) candidates per image with degenerate boxes ranked inside the top 30 000: pins the reference ORDER
# "top-max_nms by score (general.py:845-846), then obb_nms drops min(w,h) < 0.001 (nms_rotated_wrapper.py:32-39)".
# synth_pred keyword arguments and the non_max_suppression_obb arguments of the case:
PP_OVERMAX = dict(pred=dict(B=2, A=31000, nc=15, frac_obj=1.0, clusters=60, n_tiny=600),
                  kw=dict(conf_thres=0.25, iou_thres=0.45, multi_label=True, max_det=3000))
