"""Scoring + eval driver against golden values produced by the reference's own
classification.py / detection.py (tests/golden/make_eval_golden.py, committed fixture eval_metrics.json)."""
import json
import os

import pytest

from teochat_b200.eval import metrics as M

HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "eval_metrics.json")) as f:
    GOLD = json.load(f)


@pytest.mark.parametrize("name", sorted(GOLD["expected"]))
def test_metrics_match_reference(name):
    outs, exp = GOLD["cases"][name], GOLD["expected"][name]
    got = M.metrics_fn_for(name)(outs, dataset_name=name)
    assert set(got) == set(exp)
    for k, v in exp.items():
        assert got[k] == pytest.approx(v, rel=1e-12, abs=1e-12), k


def test_pixel_metrics_match_reference():
    got = M.evaluate_masks(GOLD["cases"]["xbd_loc"])
    for k, v in GOLD["evaluate_masks_xbd_loc"].items():
        assert got[k] == pytest.approx(v, rel=1e-12), k
    # the reference's positional call evaluate_masks(results, dataset) (detection.py:161) must bind the same way here
    assert M.evaluate_masks(GOLD["cases"]["xbd_loc"], "xbd_loc") == got
    # a ground-truth class outside `classes` raises like classes.index() does in the reference (detection.py:246)
    poly = "POLYGON ((0 0, 0 10, 10 10, 10 0, 0 0))"
    with pytest.raises(ValueError):
        M.region_class_f1([{"response": "destroyed", "ground_truth": "flattened", "polygon": poly}], M.DAMAGE_CLASSES)


def test_wkt_parser():
    assert M.parse_wkt_exteriors("POLYGON ((0 0, 0 2, 2 2, 2 0, 0 0), (0.5 0.5, 1 0.5, 1 1, 0.5 0.5))") == \
        [[(0.0, 0.0), (0.0, 2.0), (2.0, 2.0), (2.0, 0.0), (0.0, 0.0)]]
    multi = M.parse_wkt_exteriors("MULTIPOLYGON (((0 0, 0 1, 1 1, 0 0)), ((5 5, 5 6, 6 6, 5 5), (5.1 5.1, 5.2 5.1, 5.2 5.2, 5.1 5.1)))")
    assert [len(r) for r in multi] == [4, 4] and multi[1][0] == (5.0, 5.0)
    assert M.parse_wkt_exteriors("POLYGON EMPTY") == []
    with pytest.raises(ValueError):
        M.parse_wkt_exteriors("LINESTRING (0 0, 1 1)")


def test_unknown_dataset_and_task():
    with pytest.raises(ValueError):
        M.metrics_fn_for("imagenet")
    with pytest.raises(ValueError):
        M.detection_metrics([{"response": "a", "ground_truth": "a", "task": "captioning"}], dataset_name="xbd_loc")
    # a task without a single hit is absent from the reference's accuracy dict → KeyError there and here
    with pytest.raises(KeyError):
        M.detection_metrics([{"response": "a", "ground_truth": "b", "task": "question_answering"}], dataset_name="s2_sre_qa")


def test_eval_driver_caches_and_scores(tmp_path, monkeypatch):
    """eval(): output naming, JSON layout, cache reuse (no model load on the second call), metric dispatch."""
    from teochat_b200.eval import eval as E
    calls = {"load": 0, "infer": 0}

    def fake_load(*a, **k):
        calls["load"] += 1
        return "tok", "model", "proc"

    def fake_run_inference(dataset, model, tokenizer, processor, prompt_strategy, chronological_prefix, conv_mode, temperature,
                           max_new_tokens, batch_size=1):
        calls["infer"] += 1
        assert (model, tokenizer, processor) == ("model", "tok", "proc") and prompt_strategy == "interleave"
        return [{"response": e["conversations"][1]["value"] if i % 2 == 0 else "wrong", "ground_truth": e["conversations"][1]["value"],
                 "task": e["task"]} for i, e in enumerate(dataset)]

    monkeypatch.setattr(E, "load_model", fake_load)
    import teochat_b200.eval.inference as INF
    monkeypatch.setattr(INF, "run_inference", fake_run_inference)
    data = [{"conversations": [{"value": "q"}, {"value": f"Label {i % 3}"}], "task": "classification", "video": [], "timestamp": []}
            for i in range(10)]
    kw = dict(out_dir=str(tmp_path), prompt_strategy="interleave", chronological_prefix=True, dataset=data)
    m1 = E.eval("aid", "ckpt/teochat-synthetic-tiny", None, **kw)
    assert m1 == {"classification_accuracy": 0.5}
    out = tmp_path / "aid" / "teochat-synthetic-tiny_prompt_strategy_interleave_chronological_prefix_True.json"
    assert out.exists() and len(json.load(open(out))) == 10
    m2 = E.eval("aid", "ckpt/teochat-synthetic-tiny", None, **kw)
    assert m2 == m1 and calls == {"load": 1, "infer": 1}
    E.eval("aid", "ckpt/teochat-synthetic-tiny", None, force_rerun=True, **kw)
    assert calls == {"load": 2, "infer": 2}
    with pytest.raises(ValueError):
        E.eval("imagenet", "ckpt/teochat-synthetic-tiny", None, **kw)
