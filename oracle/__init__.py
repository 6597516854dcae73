"""CPU oracle for the caesar-yolo tiled source-finding hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under caesar_yolo_b200/ may import this package; only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it, and only as the checker
or the timed CPU baseline.

PARITY STATUS — pinned where the reference's code can execute here, unpinned for the absent third-party packages:
 * PINNED to the reference itself: tests/golden/ref_golden.json holds outputs of the UNMODIFIED reference modules
   (caesar_yolo/{utils,graph,evaluation,inference,preprocessing}.py imported from /root/reference with inert stubs for
   the absent imports; generator tests/golden/make_ref_golden.py) for generate_tiles, get_iou, get_merged_bbox, Graph,
   process_detections, make_json_results, tile neighbours, find_sources_at_edge, merge_edge_sources, MinMaxNormalizer,
   ChanResizer; tests/test_reference_golden_cpu.py requires this oracle to reproduce them exactly (and
   tests/test_reference_golden_gpu.py compares the CUDA path with the same vectors directly).
 * PINNED as glue, primitives substituted: the reference's BkgSubtractor / SigmaClipShifter / SigmaClipper /
   ZScaleTransformer / HistEqualizer / Chan3Trasformer and run.py's stage order were executed with
   astropy.stats.sigma_clipped_stats / sigma_clip, astropy.visualization.ZScaleInterval and
   skimage.exposure.equalize_hist backed by oracle/astro.py (tests/golden/ref_preproc.npz).
 * PARITY UNPINNED for the third-party algorithms themselves: astropy, scikit-image, ultralytics, fitsio, regions,
   mpi4py are absent and un-vendored (versions unpinned in the reference's requirements.txt).  The oracle restates
   their published algorithms (astropy.stats.sigma_clipping, ZScaleInterval, skimage equalize_hist, ultralytics
   LetterBox / YOLOv8 DetectionModel / Detect decode / non_max_suppression / scale_boxes) and uses
   the installed libraries directly where they exist (numpy, cv2.resize, torchvision.ops.nms, torch CPU conv2d).
   oracle/yolo11.py (YOLO11: C3k2 / C2PSA / depthwise-separable Detect) is such a restatement too.  The layer tables
   of both model families are pinned by the parameter counts and GFLOPs ultralytics publishes for all ten scales
   (tests/test_weights_cpu.py); the forward semantics on top of them are not.
   The reference's only fixture, test/galaxy0001.fits, is a known-answer input whose expected statistics match the
   survey-time hand probe (tests/golden/galaxy0001_golden.json, make_golden.py).
"""
