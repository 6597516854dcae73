"""Restatement of caesar_yolo/utils.py and caesar_yolo/graph.py pieces on the hot path (oracle; test-only)."""
import sys

import numpy as np


def get_iou(bb1, bb2):
    """caesar_yolo/utils.py:54-107.  Operands are numpy.float32 scalars in the reference
    (evaluation.py:262,302); with numpy>=2 all arithmetic stays float32."""
    bb1 = [np.float32(v) for v in bb1]
    bb2 = [np.float32(v) for v in bb2]
    assert bb1[0] < bb1[2]
    assert bb1[1] < bb1[3]
    assert bb2[0] < bb2[2]
    assert bb2[1] < bb2[3]
    x_left = max(bb1[0], bb2[0])
    y_top = max(bb1[1], bb2[1])
    x_right = min(bb1[2], bb2[2])
    y_bottom = min(bb1[3], bb2[3])
    if x_right < x_left or y_bottom < y_top:
        return np.float32(0.0)
    inter = (x_right - x_left) * (y_bottom - y_top)
    a1 = (bb1[2] - bb1[0]) * (bb1[3] - bb1[1])
    a2 = (bb2[2] - bb2[0]) * (bb2[3] - bb2[1])
    iou = inter / np.float32(a1 + a2 - inter)
    return np.float32(iou)


def get_merged_bbox(bboxes):
    """caesar_yolo/utils.py:110-119."""
    x = np.array(bboxes)
    return (min(x[:, 0]), min(x[:, 1]), max(x[:, 2]), max(x[:, 3]))


def generate_tiles(img_xmin, img_xmax, img_ymin, img_ymax, tileSizeX, tileSizeY, gridStepSizeX, gridStepSizeY):
    """caesar_yolo/utils.py:622-697.  Returns list of (xmin, xmax_excl, ymin, ymax_excl), row-major."""
    if img_xmax <= img_xmin or img_ymax <= img_ymin:
        return None
    if tileSizeX <= 0 or tileSizeY <= 0:
        return None
    if gridStepSizeX <= 0 or gridStepSizeY <= 0 or gridStepSizeX > 1 or gridStepSizeY > 1:
        return None
    Nx = img_xmax - img_xmin + 1
    Ny = img_ymax - img_ymin + 1
    if tileSizeX > Nx or tileSizeY > Ny:
        return None
    stepSizeX = int(np.round(gridStepSizeX * tileSizeX))
    stepSizeY = int(np.round(gridStepSizeY * tileSizeY))
    indexX = 0
    indexY = 0
    ix_min, ix_max, iy_min, iy_max = [], [], [], []
    while indexY <= Ny:
        offsetY = min(tileSizeY, Ny - indexY)
        ymin = indexY
        ymax = indexY + offsetY
        if ymin >= Ny or offsetY == 0:
            break
        iy_min.append(ymin)
        iy_max.append(ymax)
        indexY += stepSizeY
    while indexX <= Nx:
        offsetX = min(tileSizeX, Nx - indexX)
        xmin = indexX
        xmax = indexX + offsetX
        if xmin >= Nx or offsetX == 0:
            break
        ix_min.append(xmin)
        ix_max.append(xmax)
        indexX += stepSizeX
    tileGrid = []
    for j in range(len(iy_min)):
        for i in range(len(ix_min)):
            tileGrid.append((img_xmin + ix_min[i], img_xmin + ix_max[i], img_ymin + iy_min[j], img_ymin + iy_max[j]))
    return tileGrid


class Graph:
    """caesar_yolo/graph.py:2-41 — recursive DFS connected components; component order = ascending smallest
    vertex, member order = DFS preorder with adjacency in insertion order."""

    def __init__(self, V):
        self.V = V
        self.adj = [[] for _ in range(V)]

    def DFSUtil(self, temp, v, visited):
        visited[v] = True
        temp.append(v)
        for i in self.adj[v]:
            if not visited[i]:
                temp = self.DFSUtil(temp, i, visited)
        return temp

    def addEdge(self, v, w):
        self.adj[v].append(w)
        self.adj[w].append(v)

    def connectedComponents(self):
        # the reference recurses; raise the limit so that long chains do not abort the oracle
        old = sys.getrecursionlimit()
        sys.setrecursionlimit(max(old, self.V + 1000))
        try:
            visited = [False] * self.V
            cc = []
            for v in range(self.V):
                if not visited[v]:
                    cc.append(self.DFSUtil([], v, visited))
        finally:
            sys.setrecursionlimit(old)
        return cc
