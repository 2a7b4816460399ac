"""tools/repin.py end to end against a STAND-IN reference package (the real checkout is absent): a
throw-away package named `protoquant` whose per-token quantiser is torch.ao's decomposed op.  The tool must find it
by name, identify the knob set that reproduces it (x * (1/s), eps = 1e-5), and report "no match" with the first
differing element for an arithmetic the oracle has no knob for."""
import os
import subprocess
import sys
import textwrap

from conftest import ROOT

FAKE = '''
import torch
import torch.ao.quantization.fx._decomposed as D

def quantize_per_token(x):
    xin = x.float() if x.dtype == torch.bfloat16 else x
    s, zp = D.choose_qparams_per_token(xin, torch.int8)
    q = D.quantize_per_token(xin, s, zp, -128, 127, torch.int8)
    return q, s.to(torch.float32).flatten()

def quantize_per_channel(w):
    return quantize_per_token(w)
'''

ALIEN = '''
import torch

def quantize_per_token(x):            # round-half-AWAY: the oracle has no knob for it
    xf = x.float()
    amax = xf.abs().amax(-1, keepdim=True)
    s = torch.where(amax == 0, torch.ones_like(amax), amax / 127.0)
    r = xf / s
    q = torch.where(torch.isnan(r), torch.zeros_like(r), torch.sign(r) * torch.floor(r.abs() + 0.5)).clamp(-128, 127)
    return q.to(torch.int8), s.flatten()
'''


def _run(tmp_path, body):
    pkg = tmp_path / "protoquant"
    pkg.mkdir(exist_ok=True)
    (pkg / "__init__.py").write_text(textwrap.dedent(body))
    return subprocess.run([sys.executable, os.path.join(ROOT, "tools", "repin.py"), str(tmp_path)],
                          capture_output=True, text=True, timeout=300)


def test_repin_identifies_the_matching_knob_set(tmp_path):
    p = _run(tmp_path, FAKE)
    assert p.returncode == 0, p.stdout + p.stderr
    assert "act quantiser: protoquant:quantize_per_token" in p.stdout
    assert "PINNED: QuantSpec(scale_mode=1, eps=1e-05, qmin=-128" in p.stdout


def test_repin_reports_the_first_difference_when_no_knob_matches(tmp_path):
    p = _run(tmp_path, ALIEN)
    assert p.returncode == 1, p.stdout + p.stderr
    assert "NO knob set reproduces the reference" in p.stdout and "row" in p.stdout


def test_repin_says_so_when_the_reference_is_absent(tmp_path):
    p = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "repin.py"), str(tmp_path)], capture_output=True, text=True)
    assert p.returncode == 2 and "still absent" in p.stdout
