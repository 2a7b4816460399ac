"""Drop-in on third-party model code (the use the reference serves): Hugging Face's own LlamaMLP and BertSelfAttention
modules, converted in place by `swap_linear`.  The converted module must (a) run unchanged, (b) stay close to the float
module, (c) give the same bits with and without the shared-input fusion, (d) cost one act-quant + one GEMM per distinct
activation."""
import copy

import pytest
import torch

import protoquant_b200 as pq

pytestmark = pytest.mark.gpu

tf = pytest.importorskip("transformers")


def _llama_mlp():
    from transformers.models.llama.modeling_llama import LlamaConfig, LlamaMLP
    cfg = LlamaConfig(hidden_size=512, intermediate_size=1376, num_attention_heads=4, num_hidden_layers=1, vocab_size=64)
    return LlamaMLP(cfg)


def _bert_attention():
    from transformers.models.bert.modeling_bert import BertConfig, BertSelfAttention
    cfg = BertConfig(hidden_size=256, num_attention_heads=4, intermediate_size=512, num_hidden_layers=1)
    cfg._attn_implementation = "eager"
    return BertSelfAttention(cfg)


@pytest.mark.parametrize("make,launches", [(_llama_mlp, 4), (_bert_attention, 2)])
def test_swap_linear_on_hugging_face_modules(make, launches):
    torch.manual_seed(0)
    try:
        ref = make().to(torch.bfloat16).cuda().eval()
    except Exception as ex:                       # a transformers release with another constructor: nothing to test
        pytest.skip(f"transformers API differs: {ex!r}")
    hidden = ref.gate_proj.in_features if hasattr(ref, "gate_proj") else ref.query.in_features
    x = torch.randn(3, 40, hidden, dtype=torch.bfloat16, device="cuda")

    def run(m):
        out = m(x)
        return out[0] if isinstance(out, tuple) else out

    with torch.no_grad():
        y_ref = run(ref)
        fused = pq.swap_linear(copy.deepcopy(ref))                              # q/k/v or gate/up share one quantisation + GEMM
        plain = pq.swap_linear(copy.deepcopy(ref), fuse_shared_inputs=False)
        assert any(isinstance(m, pq.SharedInputLinear) for m in fused.modules())
        assert not any(isinstance(m, torch.nn.Linear) for m in fused.modules())
        before = pq.launch_count()
        y_fused = run(fused)
        assert pq.launch_count() - before == launches
        y_plain = run(plain)
    assert torch.equal(y_fused, y_plain)                                         # the fusion is exact
    err = (y_fused.float() - y_ref.float()).abs().max().item()
    assert err < 0.08 * y_ref.float().abs().max().item() + 1e-3                  # int8 dynamic quantisation, not a bug
    # state_dict round trip of the converted third-party module
    sd = fused.state_dict()
    again = pq.swap_linear(copy.deepcopy(ref))
    again.load_state_dict(sd)
    with torch.no_grad():
        assert torch.equal(run(again), y_fused)
