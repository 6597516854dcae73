"""Defaults dictionary with the reference's keys and values (caesar_yolo/config.py:4-59).  `scripts/run.py` mutates it
in place exactly like the reference does (scripts/run.py:311-338)."""

CONFIG = {
    # - Image resize
    'img_size': 640,
    # - Preprocessor function
    'preprocess_fcn': None,
    # - Image read options
    'image_path': '',
    'image_xmin': 0,
    'image_xmax': 0,
    'image_ymin': 0,
    'image_ymax': 0,
    # - Image parallel read options ('mpi' is accepted for compatibility; ranks come from torch.distributed)
    'mpi': None,
    'split_image_in_tiles': False,
    'tile_xsize': 256,
    'tile_ysize': 256,
    'tile_xstep': 1.0,
    'tile_ystep': 1.0,
    'max_ntasks_per_worker': 100,
    # - Source detection options
    'devices': ['cpu'],
    'use_multi_gpu': False,
    'iou_thr': 0.5,
    'merge_overlap_iou_thr_soft': 0.3,
    'merge_overlap_iou_thr_hard': 0.8,
    'score_thr': 0.7,
    # - Catalog json output options
    'save_catalog': True,
    'save_tile_catalog': False,
    'outfile_json': '',
    # - DS9 region output options
    'save_region': True,
    'save_tile_region': False,
    'outfile': '',
    # - Image output file options
    'save_img': False,
    'save_tile_img': False,
    # - Save inference plot
    'draw_plot': False,
    'draw_class_label_in_caption': True,
    'save_plot': False,
}
