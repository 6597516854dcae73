"""Catalog / region writers with the reference's file formats: `catalog_<image_id>.json`
(SFinder.write_json_results, caesar_yolo/inference.py:1196-1211), `out_<image_id>.json`
(Analyzer.write_json_results, caesar_yolo/evaluation.py:472-483) and the DS9 region text the `regions` package emits
(inference.py:1214-1287, evaluation.py:487-548; `regions` itself is not a dependency)."""
import json

CLASS_COLOR_DS9 = {  # SFinder (mosaic catalog), inference.py:334-342
    'bkg': "black", 'spurious': "red", 'compact': "blue", 'extended': "green", 'extended-multisland': "yellow",
    'flagged': "black", 'diffuse': "magenta",
}
CLASS_COLOR_DS9_ANALYZER = {  # Analyzer (single image / per-tile files), evaluation.py:108-115
    'bkg': "black", 'spurious': "red", 'compact': "blue", 'extended': "green", 'extended-multisland': "orange",
    'flagged': "magenta",
}


def sources_to_dicts(src, names):
    """cy_source array -> list of dicts with the reference's keys.  Pass-through sources keep the reference's mixed
    `edge` typing (int 0 from make_json_results, True once find_sources_at_edge fired)."""
    out = []
    for i, s in enumerate(src):
        edge = bool(int(s['flags']) & 1)
        out.append({
            "name": "S%d" % (i + 1),
            "x1": float(s['x1']), "x2": float(s['x2']), "y1": float(s['y1']), "y2": float(s['y2']),
            "class_id": int(s['cls']), "class_name": str(names[int(s['cls'])]), "score": float(s['score']),
            "edge": True if edge else 0, "merged": bool(int(s['flags']) & 2),
        })
    return out


def records_to_objs(recs, names, tag=""):
    """cy_det_record array of ONE image/tile -> Analyzer.results['objs'] (evaluation.py:418-469)."""
    objs = []
    for i, r in enumerate(recs):
        name = 'S' + str(i + 1) if tag == "" else 'S' + str(i + 1) + "_" + tag
        objs.append({"name": name, "x1": float(r['x1']), "x2": float(r['x2']), "y1": float(r['y1']),
                     "y2": float(r['y2']), "class_id": int(r['cls']), "class_name": str(names[int(r['cls'])]),
                     "score": float(r['score']), "edge": int(int(r['flags']) & 1)})
    return objs


def write_json(obj, path):
    with open(path, 'w') as fp:
        json.dump(obj, fp, indent=2, sort_keys=True)


def _fmt(v):
    return ("%.4f" % v).rstrip('0').rstrip('.') if v != int(v) else "%d" % int(v)


def write_ds9(dicts, path, merged_key=True):
    """One `box` per source: centre (x1 + dx/2, y1 + dy/2) in 1-based image coordinates, text/tag/color metadata.
    merged_key=True: SFinder.make_ds9_regions (inference.py:1214-1263, its colour map, MERGED tag);
    False: Analyzer.make_ds9_regions (evaluation.py:487-528).  Like the reference, nothing is written for an empty
    list and an unknown class name raises KeyError."""
    if not dicts:
        return
    colors = CLASS_COLOR_DS9 if merged_key else CLASS_COLOR_DS9_ANALYZER
    lines = ["# Region file format: DS9 astropy/regions", "image"]
    for d in dicts:
        dx, dy = d['x2'] - d['x1'], d['y2'] - d['y1']
        xc, yc = d['x1'] + 0.5 * dx, d['y1'] + 0.5 * dy
        tags = [d['class_name']]
        if d.get('edge'):
            tags.append('BORDER')
        if merged_key and d.get('merged'):
            tags.append('MERGED')
        meta = "text={%s} " % d['name'] + " ".join("tag={%s}" % t for t in tags)
        lines.append("box(%s,%s,%s,%s,0) # %s color=%s" % (_fmt(xc + 1), _fmt(yc + 1), _fmt(dx), _fmt(dy), meta,
                                                            colors[d['class_name']]))
    with open(path, 'w') as fp:
        fp.write("\n".join(lines) + "\n")


def write_tile_outputs(recs, tiles, tile_ids, status, names, image_id, outdir, save_json, save_regions):
    """Per-tile files of TileTask.find_sources (inference.py:218-229): `catalog_<id>_tid<N>.json` /
    `catalog_<id>_tid<N>.reg`, written by Analyzer.predict (evaluation.py:216-234) for every tile whose prediction
    ran (rejected tiles return before writing), with tile-local `edge` flags and names `S<k>_t<N>`.
    recs: cy_det_record array of this rank (tile-id order); status[i] < 0: tile tile_ids[i] was rejected."""
    import os
    import numpy as np
    order = np.argsort(recs['tile_id'], kind='stable')
    recs = recs[order]
    bounds = np.searchsorted(recs['tile_id'], np.asarray(tile_ids), side='left')
    bounds_hi = np.searchsorted(recs['tile_id'], np.asarray(tile_ids), side='right')
    written = []
    for k, tid in enumerate(tile_ids):
        if status is not None and status[k] < 0:
            continue
        objs = records_to_objs(recs[bounds[k]:bounds_hi[k]], names, tag="t" + str(int(tid)))
        base = os.path.join(outdir, 'catalog_' + str(image_id) + '_tid' + str(int(tid)))
        if save_json:
            write_json({"image_id": image_id, "objs": objs}, base + '.json')
            written.append(base + '.json')
        if save_regions and objs:
            write_ds9(objs, base + '.reg', merged_key=False)
            written.append(base + '.reg')
    return written
