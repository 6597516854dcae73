"""CPU oracle for the caesar-yolo tiled source-finding hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under caesar_yolo_b200/ may import this package; only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it, and only as the checker
or the timed CPU baseline.

PARITY UNPINNED: the reference (SKA-INAF/caesar-yolo) ships no golden vectors, no known-answer tests and cannot
be imported in this environment (astropy, scikit-image, ultralytics, fitsio, regions, mpi4py are absent and
un-vendored; versions are unpinned in the reference's requirements.txt).  The oracle therefore restates
 (1) the reference's own Python literally (file:line cited per function), and
 (2) the published algorithms of the third-party calls it makes (astropy.stats.sigma_clip /
     sigma_clipped_stats, astropy.visualization.ZScaleInterval, skimage.exposure.equalize_hist, ultralytics
     LetterBox / YOLOv8 DetectionModel / Detect decode / non_max_suppression / scale_boxes),
and uses the installed libraries directly where they exist here (numpy, cv2.resize, torchvision.ops.nms,
torch CPU conv2d).  The only reference fixture, test/galaxy0001.fits, is used as a known-answer input whose
expected statistics were derived from this restatement (tests/golden/, with the generating script).
"""
