"""CPU tests of the host-side mirror of the reference glue (prompt, tokenisation, stopping rule,
processor, config).  Expected strings/ids are written out from the reference templates
(conversation.py:252-262, inference.py:11-55, mm_utils.py:43-104)."""
import numpy as np
import pytest
import torch

from teochat_b200.config import TeoConfig
from teochat_b200.constants import IMAGE_TOKEN_INDEX
from teochat_b200.conversation import SeparatorStyle, conv_templates
from teochat_b200.eval.inference import build_prompt, extract_bboxes, replace_video_token
from teochat_b200.mm_utils import KeywordsStoppingCriteria, get_model_name_from_path, tokenizer_image_token
from teochat_b200.processor import TeoImageProcessor
from teochat_b200.tokenizer import StubTokenizer

SYSTEM = ("A chat between a curious user and an artificial intelligence assistant. "
          "The assistant gives helpful, detailed, and polite answers to the user's questions.")


def test_v1_prompt_two_style():
    conv = conv_templates["v1"].copy()
    conv.append_message(conv.roles[0], "hello <video>")
    conv.append_message(conv.roles[1], None)
    assert conv.get_prompt() == SYSTEM + " USER: hello <video> ASSISTANT:"
    assert conv.sep_style == SeparatorStyle.TWO and conv.sep2 == "</s>"
    conv.messages[-1][1] = "fine"
    conv.append_message(conv.roles[0], "again")
    assert conv.get_prompt() == SYSTEM + " USER: hello <video> ASSISTANT: fine</s>USER: again "
    assert conv_templates["v1"].messages == []          # templates are not mutated by copies


def test_replace_video_token_and_build_prompt():
    assert replace_video_token("a <video> b", ["x", "y"], None) == "a <image><image> b"
    assert replace_video_token("a <video> b", ["x", "y"], "interleave") == "a Image 1: <image>Image 2: <image> b"
    with pytest.raises(ValueError, match="Unknown prompt strategy"):
        replace_video_token("a", ["x"], "zip")
    prompt, paths, stop = build_prompt("images at times: <video> Q?", ["late", "early"], timestamps=["2021-06-01", "2019-01-31"])
    assert paths == ["early", "late"]                     # chronological sort (inference.py:45-50)
    assert "times in chronological order: Image 1: <image>Image 2: <image> Q?" in prompt
    assert stop == "</s>"
    p2, _, _ = build_prompt("images at times: <video>", ["a"], chronological_prefix=False)
    assert "times: Image 1: <image>" in p2


def test_tokenizer_image_token():
    tok = StubTokenizer()
    ids = tokenizer_image_token("hello world<image>foo<image>", tok)
    a, b = tok("hello world").input_ids, tok("foo").input_ids
    assert a[0] == tok.bos_token_id
    assert ids == a + [IMAGE_TOKEN_INDEX] + b[1:] + [IMAGE_TOKEN_INDEX]        # single BOS, -200 between chunks
    t = tokenizer_image_token("x<image>y", tok, return_tensors="pt")
    assert t.dtype == torch.long and t.tolist().count(IMAGE_TOKEN_INDEX) == 1
    with pytest.raises(ValueError):
        tokenizer_image_token("x", tok, return_tensors="np")

    class NoBos(StubTokenizer):
        def __call__(self, text, **kw):
            r = super().__call__(text)
            r.input_ids = r.input_ids[1:]
            return r
    nb = NoBos()
    assert tokenizer_image_token("a<image>b", nb) == nb("a").input_ids + [IMAGE_TOKEN_INDEX] + nb("b").input_ids


def test_keywords_stopping_criteria():
    tok = StubTokenizer()
    prompt = torch.tensor([tok("some prompt").input_ids])
    sc = KeywordsStoppingCriteria(["</s>"], tok, prompt)
    assert sc.is_eos_only(tok.eos_token_id)
    assert not sc(torch.cat([prompt, torch.tensor([[77]])], 1), None)
    assert sc(torch.cat([prompt, torch.tensor([[77, tok.eos_token_id]])], 1), None)
    sc2 = KeywordsStoppingCriteria(["stop"], tok, prompt)
    assert not sc2.is_eos_only(tok.eos_token_id)
    gen = torch.tensor([tok("go stop").input_ids[1:]])
    assert sc2(torch.cat([prompt, gen], 1), None)
    # zero new tokens: the reference slices output_ids[:, -0:] = everything, so a keyword inside the PROMPT fires (mm_utils.py:94)
    p2 = torch.tensor([tok("please stop").input_ids])
    assert KeywordsStoppingCriteria(["stop"], tok, p2)(p2, None)


def test_conversation_tuple_message_and_other_templates():
    """conversation.py:31-42: a first message that carries an image (tuple) gets its <image> tag moved to the front; the
    llava_llama_2 and plain templates build the reference's strings and stop strings."""
    from teochat_b200.conversation import conv_templates
    from teochat_b200.eval.inference import build_prompt
    c = conv_templates["v1"].copy()
    c.append_message(c.roles[0], ("What is <image> this?", object(), "Pad"))
    c.append_message(c.roles[1], None)
    assert c.get_prompt() == c.system + " USER: <image>\nWhat is  this? ASSISTANT:"
    assert c.messages[0][1][0] == "What is <image> this?"              # the stored conversation is untouched
    l2 = conv_templates["llava_llama_2"].copy()
    l2.append_message(l2.roles[0], "hi <image>")
    l2.append_message(l2.roles[1], None)
    assert l2.get_prompt() == f"[INST] <<SYS>>\n{l2.system}\n<</SYS>>\n\nhi <image> [/INST]"
    _, _, stop = build_prompt("x <video>", ["a"], conv_mode="llava_llama_2")
    assert stop == "<s>"
    prompt, _, stop = build_prompt("x <video>", ["a"], conv_mode="plain")
    assert stop == "\n" and prompt == "x Image 1: <image>\n"


def test_misc_helpers():
    assert get_model_name_from_path("/a/b/teochat-7b/") == "teochat-7b"
    assert get_model_name_from_path("/a/teochat/checkpoint-100") == "teochat_checkpoint-100"
    assert extract_bboxes("see [1, 2, 30, 40] and [5, 6, 7, 8]") == [[1, 2, 30, 40], [5, 6, 7, 8]]


def test_expand2square_and_process_images():
    """mm_utils.py:14-40 — checked against an independent numpy formulation of the padding."""
    import base64
    from io import BytesIO
    from types import SimpleNamespace

    from PIL import Image

    from teochat_b200.mm_utils import expand2square, load_image_from_base64, process_images
    rng = np.random.RandomState(1)
    fill = (122, 116, 104)
    for h, w in [(40, 40), (40, 57), (61, 30)]:
        arr = rng.randint(0, 256, (h, w, 3), dtype=np.uint8)
        sq = np.asarray(expand2square(Image.fromarray(arr), fill))
        side = max(h, w)
        want = np.empty((side, side, 3), np.uint8)
        want[:] = fill
        top, left = (side - h) // 2, (side - w) // 2
        want[top:top + h, left:left + w] = arr
        assert sq.shape == (side, side, 3) and np.array_equal(sq, want)
    p = TeoImageProcessor()
    imgs = [Image.fromarray(rng.randint(0, 256, (100, 150, 3), dtype=np.uint8)), Image.fromarray(rng.randint(0, 256, (224, 224, 3), dtype=np.uint8))]
    plain = process_images(imgs, p, SimpleNamespace())
    assert plain.shape == (2, 3, 224, 224) and torch.equal(plain[1], p.preprocess(imgs[1])["pixel_values"][0])
    padded = process_images(imgs, p, SimpleNamespace(image_aspect_ratio="pad"))
    mean_fill = tuple(int(c * 255) for c in p.image_mean)
    assert padded.shape == (2, 3, 224, 224)
    assert torch.equal(padded[0], p.preprocess(expand2square(imgs[0], mean_fill))["pixel_values"][0])
    assert not torch.equal(padded[0], plain[0]) and torch.equal(padded[1], plain[1])      # the square image is untouched
    buf = BytesIO()
    imgs[1].save(buf, format="PNG")
    assert np.array_equal(np.asarray(load_image_from_base64(base64.b64encode(buf.getvalue()))), np.asarray(imgs[1]))


def test_processor_matches_torchvision():
    from PIL import Image
    from torchvision import transforms

    from teochat_b200.constants import OPENAI_DATASET_MEAN, OPENAI_DATASET_STD
    tf = transforms.Compose([transforms.ToTensor(), transforms.Resize(224, interpolation=transforms.InterpolationMode.BICUBIC),
                             transforms.CenterCrop(224), transforms.Normalize(OPENAI_DATASET_MEAN, OPENAI_DATASET_STD)])
    p = TeoImageProcessor()
    assert p.crop_size == {"height": 224, "width": 224} and p.image_mean == OPENAI_DATASET_MEAN
    rng = np.random.RandomState(0)
    for shape in [(224, 224), (300, 256), (256, 341), (225, 400)]:
        img = Image.fromarray(rng.randint(0, 256, (shape[0], shape[1], 3), dtype=np.uint8))
        out = p.preprocess(img, return_tensors="pt")["pixel_values"]
        assert out.shape == (1, 3, 224, 224)
        assert torch.equal(out[0], tf(img))
    with pytest.raises(ValueError):
        p(images=None)
    u8 = p.to_uint8_nhwc([rng.randint(0, 256, (224, 224, 3), dtype=np.uint8)] * 2)
    assert u8.shape == (2, 224, 224, 3) and u8.dtype == torch.uint8
    with pytest.raises(ValueError):
        p.to_uint8_nhwc(rng.randint(0, 256, (100, 224, 3), dtype=np.uint8))


def test_config_derived_values():
    cfg = TeoConfig.full()
    assert cfg.vit_layers_run == 23 and cfg.tokens_per_image == 256 and cfg.vision.patch_dim == 588
    assert cfg.llama.head_dim == 128 and cfg.vision.head_dim == 64
    cfg.mm_vision_select_layer = -1
    assert cfg.vit_layers_run == 24
    cfg.mm_vision_select_layer = 30
    with pytest.raises(ValueError):
        _ = cfg.vit_layers_run
    cfg = TeoConfig.full()
    cfg.mm_vision_select_feature = "bogus"
    with pytest.raises(ValueError, match="Unexpected select feature"):
        _ = cfg.tokens_per_image
