"""Drop-in check of the command line: every flag of the reference's train_pcm_lora_sd15.py
(tests/golden/cli_flags.json, extracted by AST with tests/golden/make_cli_golden.py) exists in
`pcm_b200.train_pcm_lora_sd15` with the same default, action and type."""
import argparse
import json
import os

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "cli_flags.json")))["flags"]


@pytest.fixture(scope="module")
def parser_actions():
    from pcm_b200 import train_pcm_lora_sd15 as cli
    captured = {}
    orig = argparse.ArgumentParser.parse_args

    def grab(self, *a, **k):
        for act in self._actions:
            for s in act.option_strings:
                captured[s] = act
        return orig(self, *a, **k)

    argparse.ArgumentParser.parse_args = grab
    try:
        args = cli.parse_args([])
    finally:
        argparse.ArgumentParser.parse_args = orig
    return captured, args


def test_every_reference_flag_exists(parser_actions):
    acts, _ = parser_actions
    missing = sorted(f for f in GOLD if f not in acts)
    assert not missing, missing


def test_defaults_actions_types_match(parser_actions):
    acts, _ = parser_actions
    bad = []
    for flag, ref in GOLD.items():
        a = acts.get(flag)
        if a is None:
            continue
        if "default" in ref and not isinstance(ref["default"], dict):
            if a.default != ref["default"]:
                bad.append((flag, "default", a.default, ref["default"]))
        if ref.get("action") == "store_true" and not isinstance(a, argparse._StoreTrueAction):
            bad.append((flag, "action", type(a).__name__, "store_true"))
        if "type" in ref and ref["type"] in ("int", "float", "str"):
            if getattr(a.type, "__name__", None) != ref["type"]:
                bad.append((flag, "type", getattr(a.type, "__name__", None), ref["type"]))
        if "choices" in ref and not isinstance(ref["choices"], dict):
            if list(a.choices or []) != list(ref["choices"]):
                bad.append((flag, "choices", a.choices, ref["choices"]))
    assert not bad, bad


def test_kohya_export_keys_match_reference():
    """weights.to_kohya_keys vs the reference's get_module_kohya_state_dict executed verbatim on the
    same 278 LoRA modules (tests/golden/kohya_keys.json, make_kohya_golden.py)."""
    import torch
    from pcm_b200 import weights
    gold = json.load(open(os.path.join(HERE, "golden", "kohya_keys.json")))
    lora_sd = {}
    for m in gold["modules"]:
        lora_sd[m + ".lora_A.weight"] = torch.zeros(1)
        lora_sd[m + ".lora_B.weight"] = torch.zeros(1)
    assert len(gold["modules"]) == 278
    peft = weights.to_peft_keys(lora_sd)
    assert all(k.startswith("base_model.model.") for k in peft)
    out = weights.to_kohya_keys(lora_sd, lora_alpha=8)
    assert sorted(out.keys()) == gold["kohya_keys"]
    assert float(next(v for k, v in out.items() if k.endswith(".alpha"))) == gold["alpha"]
