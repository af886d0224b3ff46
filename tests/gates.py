"""Test helper: collect the branch decisions (ReLU / LeakyReLU gates, max-pool routing) the CUDA path took in one
step, keyed like oracle.terra_oracle.gate_tape expects, so the oracle can replay them (see the comment there)."""
import torch
import torch.nn.functional as F

from tg_b200 import plan as P


def _nchw(t):
    if t.dim() == 5:
        t = P.from_parity_split(t) if t.shape[1] == 4 else t[:, 0]
    return t.permute(0, 3, 1, 2)


def arm(G=None, D=None, crit=None):
    """Switch the engines' activation traces on (before the forward passes)."""
    if G is not None:
        G._trace = {}
    if D is not None:
        D._engine.trace = []
    if crit is not None:
        crit._vgg_engine.trace = []


def collect(G=None, D=None, crit=None) -> dict:
    gates = {}
    if G is not None:
        for k, t in G._trace.items():
            if k.endswith(".y"):
                gates[k[:-1]] = (_nchw(t) > 0).cpu()
    if D is not None:
        for p, tr in enumerate(D._engine.trace):
            for idx, t in tr.items():
                gates[f"D{p}.{idx}"] = (_nchw(t) > 0).cpu()
    if crit is not None:
        for p, tr in enumerate(crit._vgg_engine.trace):
            for idx, t in tr.items():
                y = _nchw(t).float()
                gates[f"vgg{p}.{idx}"] = (y > 0).cpu()
                if idx in (2, 7):
                    gates[f"vgg{p}.pool{idx}"] = F.max_pool2d(y, 2, 2, return_indices=True)[1].cpu()
        tr = crit._vgg_engine.trace
        if len(tr) >= 2:       # perceptual L1 (losses.py:86-89): sign(features(input) - features(target))
            gates["l1.perceptual"] = torch.sign(_nchw(tr[0][14]).float() - _nchw(tr[1][14]).float()).cpu()
    return gates


def disarm(G=None, D=None, crit=None):
    if G is not None:
        G._trace = None
    if D is not None:
        D._engine.trace = None
    if crit is not None:
        crit._vgg_engine.trace = None


def summarize(tape) -> str:
    flips = sum(v[0] for v in tape.report.values())
    total = sum(v[1] for v in tape.report.values())
    worst = max((v[2] for v in tape.report.values()), default=0.0)
    return f"{flips} of {total} branch decisions differ ({flips / max(total, 1):.1e}); largest |pre|/std among them {worst:.1e}"


def run_adversarial(G, D, criterion, real_c, masks_c, lr=2e-4, optimizers=None):
    """One iteration of train.py:179-219 with the drop-in modules, written out as the reference loop does it;
    returns outputs, losses and gradient snapshots (taken before each optimizer step)."""
    bce = torch.nn.BCEWithLogitsLoss()
    opt_G, opt_D = optimizers or (torch.optim.Adam(G.parameters(), lr=lr), torch.optim.Adam(D.parameters(), lr=lr))
    masked = real_c * masks_c
    opt_G.zero_grad()
    gen = G(masked, masks_c)
    g_loss = criterion(gen, real_c, masks_c)
    fake = D(gen)
    g_adv = bce(fake, torch.ones_like(fake))
    g_total = g_loss + g_adv
    g_total.backward()
    g_grads = {k: p.grad.detach().clone() for k, p in G.named_parameters() if p.grad is not None}
    opt_G.step()
    opt_D.zero_grad()
    real_validity = D(real_c)
    fake_validity = D(gen.detach())
    d_loss = 0.5 * (bce(real_validity, torch.ones_like(real_validity)) + bce(fake_validity, torch.zeros_like(fake_validity)))
    d_loss.backward()
    d_grads = {k: p.grad.detach().clone() for k, p in D.named_parameters() if p.grad is not None}
    opt_D.step()
    return dict(gen=gen.detach(), g_loss=g_loss.detach(), g_adv=g_adv.detach(), g_total=g_total.detach(),
                d_loss=d_loss.detach(), g_grads=g_grads, d_grads=d_grads)
