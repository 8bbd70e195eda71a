"""Generates tests/golden/chat_template_golden.json by EXECUTING the reference's own
mlx_vlm/prompt_utils.py (pure Python; loaded by path so that `import mlx` is never triggered).
Run in the build container (the GPU box has no /root/reference):  python tests/golden/make_chat_golden.py"""
import importlib.util
import json
import os

REF = "/root/reference/mlx_vlm/prompt_utils.py"
spec = importlib.util.spec_from_file_location("ref_prompt_utils", REF)
ref = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref)


class _Tok:
    """tokenizer with a template: renders a canonical string so that message structure is visible"""
    chat_template = "x"

    def apply_chat_template(self, messages, tokenize=False, add_generation_prompt=True, **kw):
        return json.dumps({"m": messages, "g": add_generation_prompt, "kw": sorted(kw)}, sort_keys=True)


class _ProcT:
    tokenizer = _Tok()


class _ProcNone:
    image_token = "<|image_pad|>"


PROCS = {"template": _ProcT(), "none": _ProcNone(), "null": None}

PROMPTS = {
    "str": "Describe this image.",
    "dict_user": {"role": "user", "content": "What is this?"},
    "dict_mm": {"role": "user", "content": [{"type": "text", "text": "look"}, {"type": "image_url", "image_url": {"url": "data:xx"}}]},
    "list_str": ["first", "second"],
    "chat": [{"role": "system", "content": "be brief"}, {"role": "user", "content": "hi"},
             {"role": "assistant", "content": "hello"}, {"role": "user", "content": "and now?"}],
    "chat_mm": [{"role": "user", "content": [{"type": "image"}, {"type": "text", "text": "a"}]},
                {"role": "assistant", "content": "ok"},
                {"role": "user", "content": [{"type": "text", "text": "b"}, {"type": "input_image"}, {"type": "image"}]}],
    "tool": [{"role": "user", "content": "call it"},
             {"role": "assistant", "content": None, "tool_calls": [{"function": {"name": "f", "arguments": "{\"a\": 1}"}}]},
             {"role": "tool", "tool_call_id": "1", "content": "42"}],
}

cases = []
for model_type in ("qwen2_vl", "qwen2_5_vl", "llava", "llava_next", "idefics2", "idefics3", "smolvlm", "some_text_model"):
    for pname, prompt in PROMPTS.items():
        for n_img in (0, 1, 3):
            for proc in ("template", "none", "null"):
                for ret in (False, True):
                    for agp in (True, False):
                        args = dict(model_type=model_type, prompt=pname, num_images=n_img, proc=proc,
                                    return_messages=ret, add_generation_prompt=agp)
                        try:
                            out = ref.apply_chat_template(PROCS[proc], {"model_type": model_type}, prompt,
                                                          add_generation_prompt=agp, return_messages=ret,
                                                          num_images=n_img)
                            cases.append({**args, "out": out})
                        except Exception as e:  # noqa
                            cases.append({**args, "error": type(e).__name__})
msgs = []
for m in ("qwen2_vl", "qwen2_5_vl", "llava", "llava_next", "idefics2", "idefics3", "smolvlm"):
    for role in ("user", "assistant", "system"):
        for n in (0, 2):
            for skip in (False, True):
                try:
                    msgs.append({"model": m, "role": role, "n": n, "skip": skip,
                                 "out": ref.get_message_json(m, "p", role, skip_image_token=skip, num_images=n, num_audios=1)})
                except Exception as e:  # noqa
                    msgs.append({"model": m, "role": role, "n": n, "skip": skip, "error": type(e).__name__})
here = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(here, "chat_template_golden.json"), "w") as f:
    json.dump({"prompts": PROMPTS, "cases": cases, "messages": msgs}, f, sort_keys=True)
print(len(cases), "cases,", len(msgs), "messages")
