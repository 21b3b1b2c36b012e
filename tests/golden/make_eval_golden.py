"""Generates tests/golden/eval_metrics.json: synthetic result dicts scored by the REFERENCE's own metric code
(/root/reference/videollava/eval/{classification,detection}.py, imported in this container only).

shapely is not installed here, so a minimal stand-in (`wkt.loads` → objects with `.exterior.coords`, MultiPolygon
iterable as in shapely 1.x) is injected before the import; everything after WKT parsing — rasterisation, confusion
matrices, per-class statistics, task dispatch — is the reference's code.  Run: python tests/golden/make_eval_golden.py
"""
import importlib.util
import json
import os
import random
import sys
import types

REF = "/root/reference/videollava/eval"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))


def install_shapely_stub():
    from teochat_b200.eval.metrics import parse_wkt_exteriors

    class Ring:
        def __init__(self, pts):
            self.coords = pts

    class Poly:
        def __init__(self, pts):
            self.exterior = Ring(pts)

    class Multi(list):
        pass

    def loads(data):
        if isinstance(data, (list, tuple)):
            return Multi(p for d in data for p in _as_list(loads(d)))
        rings = parse_wkt_exteriors(data)
        if data.strip().upper().startswith("MULTI"):
            return Multi(Poly(r) for r in rings)
        return Poly(rings[0])

    def _as_list(g):
        return list(g) if isinstance(g, Multi) else [g]

    sh = types.ModuleType("shapely")
    wkt = types.ModuleType("shapely.wkt")
    wkt.loads = loads
    sh.wkt = wkt
    sys.modules["shapely"], sys.modules["shapely.wkt"] = sh, wkt
    if "tqdm" not in sys.modules:
        try:
            import tqdm  # noqa: F401
        except ImportError:
            t = types.ModuleType("tqdm")
            t.tqdm = lambda x, **k: x
            sys.modules["tqdm"] = t


def load_ref():
    install_shapely_stub()
    pkg = types.ModuleType("videollava")
    sub = types.ModuleType("videollava.eval")
    sys.modules.setdefault("videollava_ref", pkg)
    mods = {}
    for name in ("classification", "detection"):
        spec = importlib.util.spec_from_file_location(f"videollava.eval.{name}", os.path.join(REF, f"{name}.py"))
        m = importlib.util.module_from_spec(spec)
        if name == "classification":
            saved = {k: sys.modules.get(k) for k in ("videollava", "videollava.eval", "videollava.eval.classification")}
            sys.modules["videollava"], sys.modules["videollava.eval"] = pkg, sub
            sys.modules["videollava.eval.classification"] = m
        spec.loader.exec_module(m)
        mods[name] = m
    for k, v in saved.items():
        if v is None:
            sys.modules.pop(k, None)
        else:
            sys.modules[k] = v
    return mods["classification"], mods["detection"]


def rand_box(rng):
    x1, y1 = rng.randint(0, 80), rng.randint(0, 80)
    return [x1, y1, x1 + rng.randint(3, 19), y1 + rng.randint(3, 19)]


def box_polygon(b, scale=2.56):
    x1, y1, x2, y2 = [v * scale for v in b]
    return f"POLYGON (({x1} {y1}, {x1} {y2}, {x2} {y2}, {x2} {y1}, {x1} {y1}))"


def make_cases():
    rng = random.Random(7)
    cases = {}
    labels = ["Airport", "crop field.", "Dam", "solar farm", "Port"]
    cases["fmow_high_res"] = [{"response": rng.choice(labels) + rng.choice(["", ".", "!"]), "ground_truth": rng.choice(labels),
                               "task": rng.choice(["classification", "temporal_classification"])} for _ in range(60)]
    # xBD localisation + S2 detection: boxes vs polygons, some empty on either side, a multipolygon, junk boxes
    loc = []
    for i in range(40):
        gt_boxes = [rand_box(rng) for _ in range(rng.randint(0, 3))]
        pr_boxes = [b if rng.random() < 0.5 else rand_box(rng) for b in gt_boxes] + [rand_box(rng) for _ in range(rng.randint(0, 2))]
        if len(gt_boxes) > 1:
            poly = "MULTIPOLYGON (" + ", ".join("((" + box_polygon(b)[10:-2] + "))" for b in gt_boxes) + ")"
        elif gt_boxes:
            poly = box_polygon(gt_boxes[0])
        else:
            poly = "POLYGON EMPTY"
        resp = ", ".join(str(b) for b in pr_boxes) if pr_boxes else "There are no changes."
        if i % 11 == 0 and pr_boxes:
            resp += ", [a, b, c, d]"
        loc.append({"response": resp, "ground_truth": ", ".join(str(b) for b in gt_boxes) if gt_boxes else "No buildings.",
                    "task": "change_detection_localization", "polygon": poly})
    cases["xbd_loc"] = loc
    cases["s2_det"] = [dict(o, task="change_detection_detection") for o in loc[:25]]
    dmg = ["No damage", "Minor damage", "Major damage", "Destroyed"]
    cases["xbd_dmg_cls"] = [{"response": rng.choice(dmg + ["Unknown."]) + rng.choice(["", "."]),
                             "ground_truth": rng.choice(dmg + ["Unclassified"]), "task": "change_detection_classification",
                             "polygon": box_polygon(rand_box(rng))} for _ in range(80)]
    qa = ["Yes", "No", "top left", "bottom right", "center"]
    mixed = []
    for _ in range(30):
        mixed.append({"response": rng.choice(["Yes, there is.", "no", "Yes", "It is in the top left.", "center", "Bottom right"]),
                      "ground_truth": rng.choice(qa), "task": "question_answering"})
        mixed.append({"response": rng.choice(dmg), "ground_truth": rng.choice(dmg), "task": "region_based_question_answering"})
    mixed += [dict(o, task="spatial_referring_expression") for o in loc[:20]]
    cases["xbd_sre_qa_rqa"] = mixed
    cases["s2_sre_qa"] = [o for o in mixed if o["task"] != "region_based_question_answering"]
    land = ["Residential", "Commercial", "Industrial", "Road", "Demolition", "Mega projects"]
    cases["qfabric_rqa2"] = [{"response": rng.choice(land + ["A lake"]), "ground_truth": rng.choice(land),
                              "task": "region_based_question_answering", "polygon": box_polygon(rand_box(rng))} for _ in range(70)]
    status = ["Prior construction", "Greenland", "Land cleared", "Excavation", "Materials dumped", "Construction started",
              "Construction midway", "Construction done", "Operational"]
    rq5 = [{"response": rng.choice(status), "ground_truth": rng.choice(status), "task": "region_based_temporal_question_answering",
            "polygon": box_polygon(rand_box(rng))} for _ in range(70)]
    cases["qfabric_rqa5_rtqa5"] = rq5 + cases["qfabric_rqa2"][:30]
    cases["qfabric_tre_rtqa"] = [{"response": rng.choice(["Image 1", "image 2.", "Image 3"]), "ground_truth": rng.choice(["Image 1", "Image 2", "Image 3"]),
                                  "task": rng.choice(["temporal_referring_expression", "region_based_temporal_question_answering"])}
                                 for _ in range(50)]
    return cases


def main():
    cls, det = load_ref()
    cases = make_cases()
    expected = {}
    for name, outs in cases.items():
        fn = cls.classification_metrics if name.startswith("fmow") else det.detection_metrics
        expected[name] = {k: float(v) for k, v in fn(outs, dataset_name=name).items()}
    masks = {k: float(v) for k, v in det.evaluate_masks(cases["xbd_loc"], "xbd_loc").items()}
    with open(os.path.join(HERE, "eval_metrics.json"), "w") as f:
        json.dump({"cases": cases, "expected": expected, "evaluate_masks_xbd_loc": masks,
                   "source": "reference videollava/eval/{classification,detection}.py @ /root/reference, shapely stand-in"}, f)
    for k, v in expected.items():
        print(k, v)


if __name__ == "__main__":
    main()
