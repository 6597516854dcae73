"""Restatement of caesar_yolo/evaluation.py Analyzer (predict / process_detections / make_json_results) —
oracle; test-only.  Logging and plotting are dropped; arithmetic and control flow follow the reference."""
import json

import numpy as np

from . import utils
from .utils import Graph


class Analyzer(object):
    def __init__(self, model, config):
        """evaluation.py:41-115."""
        self.model = model
        self.class_names = self.model.names
        self.config = config
        self.image = None
        self.image_id = -1
        self.image_xmin = 0
        self.image_ymin = 0
        self.bboxes_final = []
        self.class_ids_final = []
        self.scores_final = []
        self.labels_final = []
        self.results = {}
        self.obj_name_tag = ""
        self.imgsize = config['img_size']
        self.device = config['devices'][0]
        self.iou_thr = config['iou_thr']
        self.score_thr = config['score_thr']
        self.merge_overlap_iou_thr_soft = config['merge_overlap_iou_thr_soft']
        self.merge_overlap_iou_thr_hard = config['merge_overlap_iou_thr_hard']
        self.write_to_json = config.get('save_catalog', True)
        self.outfile_json = ""

    def predict(self, image, image_id='', header=None, xmin=0, ymin=0):
        """evaluation.py:128-245."""
        if image is None:
            return -1
        self.image = image
        self.image_xmin = xmin
        self.image_ymin = ymin
        if image_id:
            self.image_id = image_id
        nchans = self.image.ndim
        shp = self.image.shape
        if nchans != 3:
            cube = np.zeros((shp[0], shp[1], 3))
            cube[:, :, 0] = self.image
            cube[:, :, 1] = self.image
            cube[:, :, 2] = self.image
            self.image = cube
        dp = self.config['preprocess_fcn']
        if dp is not None:
            self.image = dp(self.image)
        if self.image is None:
            return -1
        shp = self.image.shape
        for i in range(shp[-1]):  # NB: indexes rows 0..2 (reference quirk, evaluation.py:171-176)
            if np.min(self.image[i]) == np.max(self.image[i]):
                return -1
        try:
            results = self.model(self.image, save=False, device=self.device, imgsz=self.imgsize,
                                 conf=self.score_thr, iou=self.iou_thr, visualize=False, show=False,
                                 show_labels=False, show_conf=False, show_boxes=False)
        except Exception:
            return -1
        if self.process_detections(results) < 0:
            return -1
        self.make_json_results()
        if self.write_to_json and self.outfile_json != "":
            self.write_json_results(self.outfile_json)
        return 0

    def process_detections(self, results):
        """evaluation.py:252-346."""
        bboxes_det, scores_det, labels_det, class_ids_det = [], [], [], []
        for result in results:
            bboxes = result.boxes.xyxy.cpu().numpy()
            scores = result.boxes.conf.cpu().numpy()
            cls = result.boxes.cls.cpu().numpy()
            class_labels = [self.class_names[int(item)] for item in cls]
            for i in range(len(scores)):
                if scores[i] < self.score_thr:
                    continue
                scores_det.append(scores[i])
                bboxes_det.append(bboxes[i])
                labels_det.append(class_labels[i])
                class_ids_det.append(int(cls[i]))
        self.bboxes, self.scores, self.class_ids, self.labels = bboxes_det, scores_det, class_ids_det, labels_det
        N = len(bboxes_det)
        g = Graph(N)
        for i in range(N - 1):
            for j in range(i + 1, N):
                same_class = (labels_det[i] == labels_det[j])
                iou = utils.get_iou(bboxes_det[i], bboxes_det[j])
                overlapping_soft = (iou >= self.merge_overlap_iou_thr_soft)
                overlapping_hard = (iou >= self.merge_overlap_iou_thr_hard)
                if overlapping_hard or (same_class and overlapping_soft):
                    g.addEdge(i, j)
        cc = g.connectedComponents()
        bsel, ssel, lsel, csel = [], [], [], []
        self.keep_indices = []
        for comp in cc:
            if not comp:
                continue
            score_best = 0
            index_best = -1
            for index in comp:
                if scores_det[index] > score_best:
                    score_best = scores_det[index]
                    index_best = index
            bsel.append(bboxes_det[index_best])
            lsel.append(labels_det[index_best])
            ssel.append(scores_det[index_best])
            csel.append(class_ids_det[index_best])
            self.keep_indices.append(index_best)
        self.bboxes_final, self.scores_final, self.labels_final, self.class_ids_final = bsel, ssel, lsel, csel
        return 0

    def make_json_results(self):
        """evaluation.py:418-469."""
        self.results = {"image_id": self.image_id, "objs": []}
        xmin, ymin = self.image_xmin, self.image_ymin
        ny, nx = self.image.shape[0], self.image.shape[1]
        for i in range(len(self.bboxes_final)):
            sname = 'S' + str(i + 1) if self.obj_name_tag == "" else 'S' + str(i + 1) + "_" + self.obj_name_tag
            x1, y1, x2, y2 = self.bboxes_final[i]
            x1, x2, y1, y2 = int(x1), int(x2), int(y1), int(y2)
            at_edge = False
            if x1 <= 0 or x1 >= nx - 1 or x2 <= 0 or x2 >= nx - 1:
                at_edge = True
            if y1 <= 0 or y1 >= ny - 1 or y2 <= 0 or y2 >= ny - 1:
                at_edge = True
            self.results["objs"].append({
                "name": str(sname), "x1": float(xmin + x1), "x2": float(xmin + x2), "y1": float(ymin + y1),
                "y2": float(ymin + y2), "class_id": int(self.class_ids_final[i]),
                "class_name": str(self.labels_final[i]), "score": float(self.scores_final[i]),
                "edge": int(at_edge)})

    def write_json_results(self, outfile):
        """evaluation.py:472-483."""
        if not self.results:
            return
        with open(outfile, 'w') as fp:
            json.dump(self.results, fp, indent=2, sort_keys=True)
