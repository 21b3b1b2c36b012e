"""Scoring of generated responses — the post-hoc CPU half of the reference's eval driver:
``classification_metrics`` (videollava/eval/classification.py:15-41) and ``detection_metrics``
(videollava/eval/detection.py:301-412), with the same metric names and values on the same result dicts.

Own implementation: string normalisation + per-task accuracy; binary pixel metrics from a 2×2 confusion
matrix accumulated over rasterised ground-truth polygons vs predicted boxes (256×256, box coordinates are
percentages of the image size); class-weighted per-pixel F1 for the damage / land-use tasks.  Polygons come
as WKT strings; a small POLYGON / MULTIPOLYGON parser replaces shapely (not installed here) — only the
exterior rings are rasterised, as in the reference (detection.py:143-146).
"""
from __future__ import annotations

import re
import string
from collections import Counter, defaultdict
from typing import Dict, Iterable, List, Optional, Sequence

import numpy as np

_PUNCT = str.maketrans("", "", string.punctuation)


def normalise(text: str, ignore_casing: bool = True, ignore_punctuation: bool = True) -> str:
    """classification.py:5-12"""
    if ignore_casing:
        text = text.lower()
    if ignore_punctuation:
        text = text.translate(_PUNCT)
    return text


def classification_metrics(outputs: Iterable[dict], ignore_casing=True, ignore_punctuation=True, keywords=None, **_) -> Dict[str, float]:
    """Per-task exact-match accuracy after normalisation; with ``keywords`` a response also counts when it and
    the ground truth share the first keyword found in the response order of the list (classification.py:15-41).
    Tasks without a single hit are absent from the result, like the reference's Counter-driven dict."""
    hits, totals = Counter(), Counter()
    for out in outputs:
        resp = normalise(out["response"], ignore_casing, ignore_punctuation)
        truth = normalise(out["ground_truth"], ignore_casing, ignore_punctuation)
        ok = resp == truth
        if keywords is not None:
            ok = any(k in resp and k in truth for k in keywords) or ok
        if ok:
            hits[out["task"]] += 1
        totals[out["task"]] += 1
    return {f"{task}_accuracy": n / totals[task] for task, n in hits.items()}


# ------------------------------------------------------------------------------------------ WKT → masks
_RING = re.compile(r"\(([^()]*)\)")


def parse_wkt_exteriors(wkt: str) -> List[List[tuple]]:
    """Exterior rings of a WKT POLYGON / MULTIPOLYGON (holes are dropped, detection.py:143-146)."""
    text = wkt.strip()
    head = text.split("(", 1)[0].strip().upper()
    if head.endswith("EMPTY") or "(" not in text:
        return []
    kind = head.split()[0] if head else ""
    if kind not in ("POLYGON", "MULTIPOLYGON"):
        raise ValueError(f"unsupported WKT geometry: {head!r}")

    def ring(body: str):
        pts = []
        for pair in body.split(","):
            xy = pair.split()
            if len(xy) >= 2:
                pts.append((float(xy[0]), float(xy[1])))
        return pts

    if kind == "POLYGON":
        rings = _RING.findall(text)
        return [ring(rings[0])] if rings else []
    # MULTIPOLYGON (((ext),(hole)),((ext)))
    out, depth, start = [], 0, None
    body = text[text.index("(") + 1: text.rindex(")")]
    for i, ch in enumerate(body):
        if ch == "(":
            depth += 1
            if depth == 1:
                start = i
        elif ch == ")":
            depth -= 1
            if depth == 0 and start is not None:
                rings = _RING.findall(body[start:i + 1])
                if rings:
                    out.append(ring(rings[0]))
    return out


def rasterise(rings: Sequence[Sequence[tuple]], size=(256, 256)) -> np.ndarray:
    """uint8 mask with 1 inside/on every ring (PIL polygon fill + outline, detection.py:137-158).
    Like the reference the PIL image is created with ``size`` as given, so the array is [size[1], size[0]]."""
    from PIL import Image, ImageDraw
    img = Image.new("L", tuple(size), 0)
    draw = ImageDraw.Draw(img)
    for r in rings:
        if len(r) >= 2:
            draw.polygon(list(r), outline=1, fill=1)
    return np.array(img)


_BOX = re.compile(r"\[(.*?)\]")


def boxes_from_text(text: str, width: int, height: int) -> List[List[tuple]]:
    """Predicted ``[x1, y1, x2, y2]`` boxes (percent of the image) → pixel-space rectangles (detection.py:196-209)."""
    rings = []
    for body in _BOX.findall(text):
        try:
            b = [float(t) for t in body.split(",")]
        except ValueError:
            continue
        x1, y1, x2, y2 = b[0] / 100 * width, b[1] / 100 * height, b[2] / 100 * width, b[3] / 100 * height
        rings.append([(x1, y1), (x1, y2), (x2, y2), (x2, y1), (x1, y1)])
    return rings


class PixelConfusion:
    """num_class × num_class pixel confusion matrix, rows = ground truth (detection.py:12-114)."""

    def __init__(self, num_class: int = 2):
        self.n = num_class
        self.m = np.zeros((num_class, num_class), dtype=np.int64)

    def add(self, truth: np.ndarray, pred: np.ndarray) -> None:
        if truth.shape != pred.shape:
            raise ValueError("mask shapes differ")
        keep = (truth >= 0) & (truth < self.n)
        idx = self.n * truth[keep].astype(np.int64) + pred[keep]
        self.m += np.bincount(idx, minlength=self.n * self.n).reshape(self.n, self.n)

    def binary_metrics(self) -> Dict[str, float]:
        m = self.m.astype(np.float64)
        tn, fp, fn, tp = m[0, 0], m[0, 1], m[1, 0], m[1, 1]
        total = m.sum()
        with np.errstate(divide="ignore", invalid="ignore"):
            precision = tp / (fp + tp)
            recall = tp / (fn + tp)
            f1 = 2 * recall * precision / (recall + precision)
            diag, rows, cols = np.diag(m), m.sum(1), m.sum(0)
            iou_c = diag / (rows + cols - diag + 1e-7)
            exp_acc = np.sum(cols / total * rows / total)
            obs = diag.sum() / total
            freq = rows / total
            iu = diag / (rows + cols - diag)
            return {"oa": obs, "mIoU": float(np.nanmean(iou_c)), "kappa": (obs - exp_acc) / (1 - exp_acc),
                    "fwIoU": float((freq[freq > 0] * iu[freq > 0]).sum()), "precision": precision, "recall": recall,
                    "f1": f1, "IoU": tp / (fp + fn + tp)}


def evaluate_masks(results: Iterable[dict], dataset=None, height: int = 256, width: int = 256) -> Dict[str, float]:
    """Pixel metrics of predicted boxes against the ground-truth polygons (detection.py:161-216, same positional
    signature: ``evaluate_masks(results, dataset, height=256, width=256)`` — ``dataset`` only labels the reference's
    progress bar).  A sample whose ground truth / response has no ``[`` contributes an empty mask on that side."""
    conf = PixelConfusion(2)
    for r in results:
        gt = rasterise(parse_wkt_exteriors(r["polygon"]), (height, width)) if "[" in r["ground_truth"] \
            else np.zeros((height, width), dtype=np.uint8)
        pr = rasterise(boxes_from_text(r["response"], width, height), (height, width)) if "[" in r["response"] \
            else np.zeros((height, width), dtype=np.uint8)
        conf.add(gt, pr)
    return conf.binary_metrics()


def region_class_f1(outputs: Iterable[dict], classes: Sequence[str], skip_classes: Sequence[str] = (), height=256, width=256,
                    ignore_casing=True, ignore_punctuation=True) -> Dict[str, float]:
    """Per-pixel class F1 over the regions' polygons (detection.py:219-298): every example paints its polygon with
    the predicted and the true class; returns the plain mean, the prevalence-weighted and the inverse-prevalence-
    weighted F1 over ``classes``."""
    stats = defaultdict(lambda: {"tp": 0, "fp": 0, "fn": 0, "count": 0})
    for out in outputs:
        pred = normalise(out["response"], ignore_casing, ignore_punctuation)
        truth = normalise(out["ground_truth"], ignore_casing, ignore_punctuation)
        if truth in skip_classes:
            continue
        area = int((rasterise(parse_wkt_exteriors(out["polygon"]), (height, width)) > 0).sum())
        if pred in classes:
            if truth not in classes:        # detection.py:246: classes.index(ground_truth_class) raises here
                raise ValueError(f"{truth!r} is not in list")
            # the region is painted with one predicted and one true label: all-or-nothing per example
            tp = area if pred == truth else 0
            stats[pred]["tp"] += tp
            stats[pred]["fp"] += area - tp
            stats[truth]["fn"] += area - tp
        # an out-of-vocabulary prediction adds no false negatives in the reference either (its mask is still empty there)
        stats[truth]["count"] += area
    total = sum(s["count"] for s in stats.values())
    f1s, w_f1, inv_sum, inv_w = {}, 0.0, 0.0, 0.0
    for c in classes:
        tp, fp, fn = stats[c]["tp"], stats[c]["fp"], stats[c]["fn"]
        prec = tp / (tp + fp) if tp + fp else 0.0
        rec = tp / (tp + fn) if tp + fn else 0.0
        f1 = 2 * prec * rec / (prec + rec) if prec + rec else 0.0
        f1s[c] = f1
        prev = stats[c]["count"] / total if total else 0.0
        w_f1 += f1 * prev
        if prev:
            inv_sum += f1 / prev
            inv_w += 1 / prev
    return {"f1": float(np.mean(list(f1s.values()))), "w_f1": w_f1, "inv_w_f1": inv_sum / inv_w if inv_w else 0.0}


DAMAGE_CLASSES = ["no damage", "minor damage", "major damage", "destroyed"]
LANDUSE_CLASSES = ["residential", "commercial", "industrial", "road", "demolition", "mega projects"]
STATUS_CLASSES = ["prior construction", "greenland", "land cleared", "excavation", "materials dumped", "construction started",
                  "construction midway", "construction done", "operational"]
QA_KEYWORDS = ["yes", "no", "top left", "top center", "top right", "center left", "center", "center right", "bottom left",
               "bottom center", "bottom right"]


def detection_metrics(outputs: Iterable[dict], dataset_name: str, ignore_casing=True, ignore_punctuation=True) -> Dict[str, float]:
    """Task → metric dispatch of detection.py:301-412 (same keys: ``<task>_f1`` / ``<task>_accuracy``)."""
    by_task: Dict[str, List[dict]] = defaultdict(list)
    for out in outputs:
        by_task[out["task"]].append(out)
    kw = dict(ignore_casing=ignore_casing, ignore_punctuation=ignore_punctuation)

    def acc(task, keywords=None):
        return classification_metrics(by_task[task], keywords=keywords, **kw)[f"{task}_accuracy"]

    def need(task, *names):
        if dataset_name not in names:
            raise ValueError(f"Unsupported task {task} for dataset {dataset_name}")

    res: Dict[str, float] = {}
    for task, items in by_task.items():
        if "xbd" in dataset_name:
            if task == "change_detection_classification":
                need(task, "xbd_dmg_cls")
                res[f"{task}_f1"] = region_class_f1(items, DAMAGE_CLASSES, skip_classes=["unclassified"], **kw)["inv_w_f1"]
            elif task == "change_detection_localization":
                res[f"{task}_f1"] = evaluate_masks(items)["f1"]
            elif task == "spatial_referring_expression":
                need(task, "xbd_sre_qa_rqa")
                res[f"{task}_f1"] = evaluate_masks(items)["f1"]
            elif task == "region_based_question_answering":
                need(task, "xbd_sre_qa_rqa")
                res[f"{task}_accuracy"] = acc(task)
            elif task == "question_answering":
                need(task, "xbd_sre_qa_rqa")
                res[f"{task}_accuracy"] = acc(task, QA_KEYWORDS)
            else:
                raise ValueError(f"Unsupported task {task} for dataset {dataset_name}")
        elif "s2" in dataset_name:
            if task == "change_detection_detection" and dataset_name == "s2_det":
                res[f"{task}_f1"] = evaluate_masks(items)["f1"]
            elif task == "region_based_question_answering":
                need(task, "s2_rqa")
                res[f"{task}_accuracy"] = acc(task)
            elif task == "spatial_referring_expression":
                need(task, "s2_sre_qa")
                res[f"{task}_f1"] = evaluate_masks(items)["f1"]
            elif task == "question_answering":
                need(task, "s2_sre_qa")
                res[f"{task}_accuracy"] = acc(task)
            else:
                raise ValueError(f"Unsupported task {task} for dataset {dataset_name}")
        elif "qfabric" in dataset_name:
            if task == "region_based_question_answering":
                res[f"{task}_f1"] = region_class_f1(items, LANDUSE_CLASSES, **kw)["w_f1"]
            elif task == "region_based_temporal_question_answering":
                if dataset_name == "qfabric_tre_rtqa":
                    res[f"{task}_accuracy"] = acc(task)
                elif dataset_name == "qfabric_rqa5_rtqa5":
                    res[f"{task}_f1"] = region_class_f1(items, STATUS_CLASSES, **kw)["w_f1"]
                else:
                    raise ValueError(f"Unsupported dataset {dataset_name} for task {task}")
            elif task == "temporal_referring_expression":
                need(task, "qfabric_tre_rtqa")
                res[f"{task}_accuracy"] = acc(task)
            else:
                raise ValueError(f"Unsupported task: {task} for dataset {dataset_name}")
        else:
            raise ValueError(f"Unsupported dataset: {dataset_name}")
    return res


CLASSIFICATION_DATASETS = ("fmow_high_res", "fmow_low_res", "abcd", "cdvqa", "aid", "ucm", "lrben", "hrben")
DETECTION_DATASETS = ("xbd_loc", "xbd_dmg_cls", "s2_det", "xbd_sre_qa_rqa", "s2_sre_qa", "s2_rqa", "qfabric_rqa2",
                      "qfabric_rqa5_rtqa5", "qfabric_tre_rtqa")
# dataset name → split suffix of jirvin16/TEOChatlas (eval.py:90-108)
HF_SPLITS = {
    "fmow_high_res": "fMoW_High_Res", "fmow_low_res": "fMoW_Low_Res", "abcd": "ABCD", "cdvqa": "CDVQA", "aid": "AID",
    "ucm": "UCMerced", "lrben": "LRBEN", "hrben": "HRBEN", "xbd_loc": "xBD_Change_Detection_Localization",
    "xbd_dmg_cls": "xBD_Change_Detection_Classification", "s2_det": "S2Looking_Change_Detection",
    "xbd_sre_qa_rqa": "xBD_SRE_QA_RQA", "s2_sre_qa": "S2Looking_SRE_QA", "s2_rqa": "S2Looking_RQA",
    "qfabric_rqa2": "QFabric_RQA2", "qfabric_rqa5_rtqa5": "QFabric_RQA5_RTQA5", "qfabric_tre_rtqa": "QFabric_TRE_RTQA",
}


def metrics_fn_for(dataset_name: str):
    if dataset_name in CLASSIFICATION_DATASETS:
        return classification_metrics
    if dataset_name in DETECTION_DATASETS:
        return detection_metrics
    raise ValueError(f"Unsupported dataset: {dataset_name}")
