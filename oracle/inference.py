"""Restatement of caesar_yolo/inference.py (TileTask, SFinder) — oracle; test-only.  Emulates "mpi4py present,
nproc=1" (the reference's tiled path yields an empty catalog without MPI).  Logging dropped."""
import json
import os

from . import fits_min, utils
from .evaluation import Analyzer
from .utils import Graph


class MergedSourceInfo(object):
    def __init__(self, sindex, tindex):
        self.sindex = sindex
        self.tindex = tindex


class TileTask(object):
    """inference.py:57-275."""

    def __init__(self, tile_coords, model, config):
        self.model = model
        self.config = config
        self.ix_min, self.ix_max, self.iy_min, self.iy_max = tile_coords
        self.wid = -1
        self.tid = 0
        self.sname_tag = ""
        self.neighborTaskId = []
        self.neighborTaskIndex = []
        self.neighborWorkerId = []
        self.image_id = os.path.splitext(os.path.basename(os.path.abspath(config['image_path'])))[0]
        self.det_sources = {}

    def set_task_id(self, tid):
        self.tid = tid
        self.sname_tag = "t" + str(tid)

    def is_task_tile_adjacent(self, a):
        ax = (self.ix_max == a.ix_min - 1 or self.ix_min == a.ix_max + 1 or
              (self.ix_min == a.ix_min and self.ix_max == a.ix_max))
        ay = (self.iy_max == a.iy_min - 1 or self.iy_min == a.iy_max + 1 or
              (self.iy_min == a.iy_min and self.iy_max == a.iy_max))
        return ax and ay

    def is_task_tile_overlapping(self, a):
        if self.ix_max < a.ix_min:
            return False
        if self.ix_min > a.ix_max:
            return False
        if self.iy_max < a.iy_min:
            return False
        if self.iy_min > a.iy_max:
            return False
        return True

    def is_task_tile_neighbor(self, a):
        return self.is_task_tile_adjacent(a) or self.is_task_tile_overlapping(a)

    def add_neighbor_info(self, tid, tindex, wid):
        self.neighborTaskId.append(tid)
        self.neighborTaskIndex.append(tindex)
        self.neighborWorkerId.append(wid)

    def find_sources(self):
        res = fits_min.read_fits_crop(self.config['image_path'], self.ix_min, self.ix_max, self.iy_min, self.iy_max)
        if res is None:
            return -1
        imgdata, _ = res
        analyzer = Analyzer(self.model, self.config)
        analyzer.obj_name_tag = self.sname_tag
        analyzer.write_to_json = False
        if analyzer.predict(imgdata, self.image_id, xmin=self.ix_min, ymin=self.iy_min) < 0:
            return -1
        if not analyzer.bboxes_final:
            return 0
        self.det_sources = analyzer.results
        self.det_sources["workerId"] = self.wid
        self.det_sources["tileId"] = self.tid
        self.det_sources["neighborTileIds"] = self.neighborTaskId
        self.det_sources["xmin"] = self.ix_min
        self.det_sources["xmax"] = self.ix_max
        self.det_sources["ymin"] = self.iy_min
        self.det_sources["ymax"] = self.iy_max
        return 0


class SFinder(object):
    """inference.py:280-1290, nproc=1."""

    def __init__(self, model, config):
        self.config = config
        self.model = model
        self.nproc = 1
        self.procId = 0
        self.tile_sources = {"sources": []}
        self.sources = {"sources": []}
        self.tasks_per_worker = []
        self.outdir = config.get('outdir', '.')

    def set_img_size_params(self):
        """inference.py:354-477 (FITS only; sub-image mode sets xmin.. but keeps the reference's nx/ny)."""
        hdr, _ = fits_min.read_header(self.config['image_path'])
        self.header = hdr
        xmin, xmax = self.config['image_xmin'], self.config['image_xmax']
        ymin, ymax = self.config['image_ymin'], self.config['image_ymax']
        if xmin >= 0 and xmax >= 0 and ymin >= 0 and ymax >= 0 and not (xmin == xmax == ymin == ymax == 0):
            self.xmin, self.xmax, self.ymin, self.ymax = xmin, xmax, ymin, ymax
            self.nx, self.ny = xmax - xmin + 1, ymax - ymin + 1
        else:
            self.nx, self.ny = hdr['NAXIS1'], hdr['NAXIS2']
            self.xmin, self.xmax, self.ymin, self.ymax = 0, self.nx - 1, 0, self.ny - 1
        self.tileSizeX, self.tileSizeY = self.nx, self.ny
        self.tileStepSizeX = self.tileStepSizeY = 1
        if self.config['split_image_in_tiles']:
            self.tileSizeX, self.tileSizeY = self.config['tile_xsize'], self.config['tile_ysize']
            self.tileStepSizeX, self.tileStepSizeY = self.config['tile_xstep'], self.config['tile_ystep']
        self.image_id = os.path.splitext(os.path.basename(os.path.abspath(self.config['image_path'])))[0]
        return 0

    def run(self):
        """inference.py:485-552 — single image."""
        if self.set_img_size_params() < 0:
            return -1
        res = fits_min.read_fits_crop(self.config['image_path'], self.config['image_xmin'],
                                      self.config['image_xmax'], self.config['image_ymin'],
                                      self.config['image_ymax'])
        if res is None:
            return -1
        analyzer = Analyzer(self.model, self.config)
        analyzer.outfile_json = os.path.join(self.outdir, 'out_' + str(self.image_id) + '.json')
        if analyzer.predict(image=res[0], image_id=self.image_id) < 0:
            return -1
        self.analyzer = analyzer
        return 0

    def create_tile_tasks(self):
        """inference.py:992-1162 (nproc=1)."""
        grid = utils.generate_tiles(self.xmin, self.xmax, self.ymin, self.ymax, self.tileSizeX, self.tileSizeY,
                                    self.tileStepSizeX, self.tileStepSizeY)
        if grid is None:
            return -1
        tasks = []
        for i, coords in enumerate(grid):
            t = TileTask(coords, self.model, self.config)
            t.wid = 0
            t.set_task_id(i)
            tasks.append(t)
        self.tasks_per_worker = [tasks]
        n = len(tasks)
        for j in range(n):
            for k in range(j + 1, n):
                if tasks[j].is_task_tile_neighbor(tasks[k]):
                    tasks[j].add_neighbor_info(tasks[k].tid, k, 0)
                    tasks[k].add_neighbor_info(tasks[j].tid, j, 0)
        if n > self.config['max_ntasks_per_worker']:
            return -1
        return 0

    def find_sources_at_edge(self, tindex):
        """inference.py:663-726."""
        tile = self.tasks_per_worker[0][tindex]
        data = tile.det_sources
        if not data or not data["objs"]:
            return
        xmin, xmax, ymin, ymax = tile.ix_min, tile.ix_max, tile.iy_min, tile.iy_max
        for i, s in enumerate(data["objs"]):
            x1, x2, y1, y2 = s["x1"], s["x2"], s["y1"], s["y2"]
            if (x1 == xmin or x2 == xmax) or (y1 == ymin or y2 == ymax):
                data["objs"][i]["edge"] = True
                continue
            for j in range(len(tile.neighborWorkerId)):
                n = self.tasks_per_worker[tile.neighborWorkerId[j]][tile.neighborTaskIndex[j]]
                if (x2 < n.ix_min) or (x1 > n.ix_max) or (y2 < n.iy_min) or (y1 > n.iy_max):
                    continue
                data["objs"][i]["edge"] = True
                break

    def merge_edge_sources(self):
        """inference.py:731-931."""
        to_merge = []
        self.sources["sources"] = []
        ts = self.tile_sources["sources"]
        for ti in range(len(ts)):
            objs = ts[ti]["objs"]
            for j in range(len(objs)):
                if not objs[j]["edge"]:
                    objs[j]["merged"] = False
                    self.sources["sources"].append(objs[j])
                    continue
                to_merge.append(MergedSourceInfo(j, ti))
        N = len(to_merge)
        g = Graph(N)
        for i in range(N):
            si = ts[to_merge[i].tindex]["objs"][to_merge[i].sindex]
            neigh = ts[to_merge[i].tindex]["neighborTileIds"]
            xmin, xmax, ymin, ymax = si["x1"], si["x2"], si["y1"], si["y2"]
            for j in range(i + 1, N):
                sj = ts[to_merge[j].tindex]["objs"][to_merge[j].sindex]
                tid_j = ts[to_merge[j].tindex]["tileId"]
                if tid_j not in neigh:
                    continue
                if (xmax < sj["x1"]) or (xmin > sj["x2"]) or (ymax < sj["y1"]) or (ymin > sj["y2"]):
                    continue
                g.addEdge(i, j)
        cc = g.connectedComponents()
        for i in range(len(cc)):
            if not cc[i]:
                continue
            sname_merged = "S" + str(i + 1) + "_merged"
            if len(cc[i]) == 1:
                m = to_merge[cc[i][0]]
                s = ts[m.tindex]["objs"][m.sindex]
                s["name"] = sname_merged
                s["merged"] = False
                self.sources["sources"].append(s)
            else:
                index_largest = -1
                area_largest = -1
                bboxes = []
                for index in cc[i]:
                    m = to_merge[index]
                    s = ts[m.tindex]["objs"][m.sindex]
                    bbox = (s["x1"], s["y1"], s["x2"], s["y2"])
                    area = (s["x2"] - s["x1"]) * (s["y2"] - s["y1"])
                    if area > area_largest:
                        area_largest = area
                        index_largest = index
                    bboxes.append(bbox)
                m = to_merge[index_largest]
                sl = ts[m.tindex]["objs"][m.sindex]
                bm = utils.get_merged_bbox(bboxes)
                self.sources["sources"].append({
                    "name": str(sname_merged), "x1": float(bm[0]), "x2": float(bm[2]), "y1": float(bm[1]),
                    "y2": float(bm[3]), "edge": True, "merged": True, "score": sl["score"],
                    "class_name": sl["class_name"], "class_id": sl["class_id"]})
        for i in range(len(self.sources["sources"])):
            self.sources["sources"][i]["name"] = "S" + str(i + 1)
        return 0

    def run_parallel(self):
        """inference.py:578-658 with nproc=1."""
        if self.set_img_size_params() < 0:
            return -1
        if self.create_tile_tasks() < 0:
            return -1
        tasks = self.tasks_per_worker[0]
        for j in range(len(tasks)):
            if tasks[j].find_sources() < 0:
                continue
            self.find_sources_at_edge(j)
        self.tile_sources = {"sources": [t.det_sources for t in tasks if t.det_sources]}
        self.merge_edge_sources()
        self.save()
        return 0

    def save(self):
        """inference.py:1167-1211 (catalog json only; DS9 writer needs the `regions` package)."""
        out = os.path.join(self.outdir, 'catalog_' + str(self.image_id) + '.json')
        with open(out, 'w') as fp:
            json.dump(self.sources, fp, indent=2, sort_keys=True)
        self.outfile_json = out
