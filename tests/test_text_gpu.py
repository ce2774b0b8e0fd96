"""The text encoder on this library's kernels (tubedetr_b200/text.py) vs the HF RobertaModel it replaces (reference
models/transformer.py:130-135, 250-263 calls the HF module; fp32 eager there), same weights, padded captions."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _roberta(seed=0):
    from transformers import RobertaConfig, RobertaModel
    torch.manual_seed(seed)
    cfg = RobertaConfig(vocab_size=50265, max_position_embeddings=514, type_vocab_size=1, pad_token_id=1, bos_token_id=0, eos_token_id=2,
                        layer_norm_eps=1e-5)
    m = RobertaModel(cfg)
    with torch.no_grad():          # non-trivial LayerNorm parameters / biases (HF initialises them to 1 / 0)
        for n, p in m.named_parameters():
            if "LayerNorm.weight" in n:
                p.add_(torch.randn_like(p) * 0.05)
            elif n.endswith("bias"):
                p.add_(torch.randn_like(p) * 0.02)
    return m.cuda()


@pytest.mark.parametrize("B,L", [(1, 20), (3, 11), (2, 37)])
def test_text_encoder_matches_hf_roberta(B, L):
    from tubedetr_b200 import text
    m = _roberta().eval()
    g = torch.Generator().manual_seed(5)
    ids = torch.randint(3, 50000, (B, L), generator=g)
    am = torch.ones(B, L, dtype=torch.long)
    for b in range(1, B):
        n = L - 2 * b
        ids[b, n:] = 1
        am[b, n:] = 0
    ids[:, 0] = 0
    ids, am = ids.cuda(), am.cuda()
    assert text.supported(m, ids)
    ref = m(input_ids=ids, attention_mask=am).last_hidden_state                      # fp32 library path
    x32, xb = text.roberta_forward(m, ids, am, training=False)
    got = x32.view(B, L, 768)
    valid = am.bool()
    err = (got - ref)[valid].abs().max().item()
    assert err <= 2e-2 * ref[valid].abs().max().item(), err
    assert (got - ref)[valid].abs().mean().item() <= 4e-3 * ref[valid].abs().mean().item()
    # gradients of a random linear functional of the valid rows
    w = torch.randn(B, L, 768, generator=g).cuda() * valid[..., None]
    names = ["embeddings.word_embeddings.weight", "embeddings.LayerNorm.weight", "encoder.layer.0.attention.self.query.weight",
             "encoder.layer.3.attention.self.query.bias", "encoder.layer.5.attention.output.dense.weight", "encoder.layer.7.intermediate.dense.weight",
             "encoder.layer.7.intermediate.dense.bias", "encoder.layer.11.output.dense.weight", "encoder.layer.11.output.LayerNorm.bias",
             "encoder.layer.2.attention.self.value.weight"]
    params = dict(m.named_parameters())
    gr = torch.autograd.grad((ref * w).sum(), [params[n] for n in names])
    go = torch.autograd.grad((got * w).sum(), [params[n] for n in names])
    for n, a, c in zip(names, go, gr):
        cos = (a * c).sum() / (a.norm() * c.norm() + 1e-30)
        assert cos > 0.995 and abs(a.norm().item() / (c.norm().item() + 1e-30) - 1) < 0.03, (n, cos.item(), a.norm().item(), c.norm().item())


def test_text_encoder_train_mode_dropout_and_small_kernels():
    """train(): dropouts live (two passes differ), finite gradients; GELU / skinny weight-gradient kernels vs torch"""
    from tubedetr_b200 import kernels as K
    from tubedetr_b200 import text
    m = _roberta(1).train()
    ids = torch.randint(3, 50000, (2, 16)).cuda()
    am = torch.ones(2, 16, dtype=torch.long).cuda()
    a, _ = text.roberta_forward(m, ids, am, training=True)
    a.float().square().sum().backward()
    assert all(torch.isfinite(p.grad).all() for n, p in m.named_parameters() if p.grad is not None)
    assert sum(p.grad is not None for p in m.parameters()) >= 195
    from tubedetr_b200 import ops
    ops.advance_dropout_seed(ids.device)
    b, _ = text.roberta_forward(m, ids, am, training=True)
    assert (a - b).abs().max() > 1e-3
    x = torch.randn(20, 3072, device="cuda").bfloat16()
    torch.testing.assert_close(K.gelu_fwd(x, torch.empty_like(x)).float(), torch.nn.functional.gelu(x.float()).bfloat16().float(), atol=2e-2, rtol=2e-2)
    xf = x.float().requires_grad_(True)
    dy = torch.randn(20, 3072, device="cuda").bfloat16()
    (torch.nn.functional.gelu(xf) * dy.float()).sum().backward()
    torch.testing.assert_close(K.gelu_bwd(dy, x, torch.empty_like(x)).float(), xf.grad, atol=3e-2, rtol=3e-2)
    for R, N, Kd in ((20, 768, 768), (37, 3072, 768), (5, 768, 3072)):
        dyb, xb = torch.randn(R, N, device="cuda").bfloat16(), torch.randn(R, Kd, device="cuda").bfloat16()
        dW, db = torch.empty(N, Kd, device="cuda"), torch.empty(N, device="cuda")
        K.skinny_wgrad(dyb, xb, dW, db)
        torch.testing.assert_close(dW, dyb.float().t() @ xb.float(), atol=1e-3, rtol=1e-4)
        torch.testing.assert_close(db, dyb.float().sum(0), atol=1e-4, rtol=1e-5)
