"""Argument sets shared by tests/golden/make_postprocess_golden.py and the tests."""
PP_CFGS = {"multi": dict(conf_thres=0.25, iou_thres=0.45, multi_label=True, max_det=1500),
           "best": dict(conf_thres=0.3, iou_thres=0.4, multi_label=False, max_det=1000),
           "cls": dict(conf_thres=0.25, iou_thres=0.45, multi_label=True, classes=[1, 5, 9], max_det=50),
           "agn": dict(conf_thres=0.28, iou_thres=0.2, multi_label=True, agnostic=True, max_det=1500),
           "lowconf": dict(conf_thres=0.001, iou_thres=0.6, multi_label=True, max_det=300)}
PP_ANCHORS = {"lowconf": 100}
