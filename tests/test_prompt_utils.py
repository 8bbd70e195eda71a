"""apply_chat_template / get_message_json (mlx_vlm/prompt_utils.py:555-995) against the reference's
own module executed on 1008 prompt shapes x processors (tests/golden/make_chat_golden.py)."""
import json
import os

from mlx_vlm_b200 import apply_chat_template, get_message_json

HERE = os.path.dirname(os.path.abspath(__file__))


class _Tok:
    chat_template = "x"

    def apply_chat_template(self, messages, tokenize=False, add_generation_prompt=True, **kw):
        return json.dumps({"m": messages, "g": add_generation_prompt, "kw": sorted(kw)}, sort_keys=True)


class _ProcT:
    tokenizer = _Tok()


class _ProcNone:
    image_token = "<|image_pad|>"


def test_chat_template_matches_reference_module():
    with open(os.path.join(HERE, "golden", "chat_template_golden.json")) as f:
        G = json.load(f)
    procs = {"template": _ProcT(), "none": _ProcNone(), "null": None}
    n_ok = 0
    for case in G["cases"]:
        prompt = G["prompts"][case["prompt"]]
        try:
            out = apply_chat_template(procs[case["proc"]], {"model_type": case["model_type"]}, prompt,
                                      add_generation_prompt=case["add_generation_prompt"],
                                      return_messages=case["return_messages"], num_images=case["num_images"])
            got = {"out": json.loads(json.dumps(out))}
        except Exception as e:  # noqa
            got = {"error": type(e).__name__}
        want = {k: case[k] for k in ("out", "error") if k in case}
        assert got == want, (case, got)
        n_ok += 1
    assert n_ok == len(G["cases"]) >= 1000
    for m in G["messages"]:
        try:
            got = {"out": json.loads(json.dumps(get_message_json(m["model"], "p", m["role"], skip_image_token=m["skip"],
                                                                 num_images=m["n"], num_audios=1)))}
        except Exception as e:  # noqa
            got = {"error": type(e).__name__}
        assert got == {k: m[k] for k in ("out", "error") if k in m}, m
