"""Train-step bodies of the reference's two training loops, driven through the drop-in modules.

`AdversarialStep.run` is the body of the hot loop in mvp_gan/src/train.py:179-225 (generator step,
discriminator step, BCEWithLogits, Adam x2); `HumanGuidedStep.run` is the body of
mvp_gan/src/training/human_guided_trainer.py:101-153. They exist so that bench.py, smoke() and the
parity tests exercise exactly the sequence the reference loops execute, with two additions the
reference does not have: the data-parallel gradient exchange (tg_b200.ddp) and — optionally — skipping
the discriminator weight gradients of the generator step, which train.py:210 zeroes before they are
ever used (output-equivalent; reported separately in bench.py).

Everything stays on the device: no `.item()` is called here (the reference syncs 4+ times per batch,
train.py:222-225); callers read the returned loss tensors when they want them.
"""
from __future__ import annotations

from contextlib import contextmanager
from typing import Dict, Optional

import torch

from .ddp import BucketedGradReducer
from .functional import BCEWithLogitsLoss


@contextmanager
def _frozen(module, enabled: bool):
    if not enabled:
        yield
        return
    flags = [(p, p.requires_grad) for p in module.parameters()]
    for p, _ in flags:
        p.requires_grad_(False)
    try:
        yield
    finally:
        for p, f in flags:
            p.requires_grad_(f)


class AdversarialStep:
    def __init__(self, generator, discriminator, criterion, optimizer_G, optimizer_D,
                 reducer: Optional[BucketedGradReducer] = None, skip_discarded_d_wgrad: bool = True):
        self.G, self.D, self.criterion = generator, discriminator, criterion
        self.opt_G, self.opt_D = optimizer_G, optimizer_D
        self.adversarial_loss = BCEWithLogitsLoss()                   # train.py:115 (tg_bce_logits_fwd/bwd)
        self.reducer = reducer
        self.skip_discarded_d_wgrad = skip_discarded_d_wgrad

    def run(self, real_imgs: torch.Tensor, masks: torch.Tensor) -> Dict[str, torch.Tensor]:
        masked_imgs = real_imgs * masks                                 # :181
        # ---- generator ---- :184-207
        self.opt_G.zero_grad()
        gen_imgs = self.G(masked_imgs, masks)
        g_loss = self.criterion(gen_imgs, real_imgs, masks)
        with _frozen(self.D, self.skip_discarded_d_wgrad):
            fake_validity = self.D(gen_imgs)
        g_adv_loss = self.adversarial_loss(fake_validity, 1.0)          # target = ones_like (:203)
        g_total_loss = g_loss + g_adv_loss
        g_total_loss.backward()
        if self.reducer is not None:
            self.reducer.finish()
        self.opt_G.step()
        # ---- discriminator ---- :210-219
        self.opt_D.zero_grad()
        real_validity = self.D(real_imgs)
        fake_validity = self.D(gen_imgs.detach())
        real_loss = self.adversarial_loss(real_validity, 1.0)          # :213-216
        fake_loss = self.adversarial_loss(fake_validity, 0.0)
        d_loss = 0.5 * (real_loss + fake_loss)
        d_loss.backward()
        if self.reducer is not None:
            self.reducer.finish()
        self.opt_D.step()
        return dict(g_total_loss=g_total_loss.detach(), g_loss=g_loss.detach(), g_adv_loss=g_adv_loss.detach(),
                    d_loss=d_loss.detach(), real_loss=real_loss.detach(), fake_loss=fake_loss.detach(),
                    gen_imgs=gen_imgs.detach())


class HumanGuidedStep:
    def __init__(self, generator, criterion, optimizer, reducer: Optional[BucketedGradReducer] = None):
        self.G, self.criterion, self.opt, self.reducer = generator, criterion, optimizer, reducer

    def run(self, images, masks, human_masks=None) -> Dict[str, torch.Tensor]:
        masked_images = images * masks                                  # human_guided_trainer.py:112
        generated = self.G(masked_images, masks)                        # :115
        human_feedback = {'mask': human_masks} if human_masks is not None else None
        loss = self.criterion(generated, images, masks, human_feedback)  # :121
        self.opt.zero_grad()                                            # :151-153
        loss.backward()
        if self.reducer is not None:
            self.reducer.finish()
        self.opt.step()
        return dict(loss=loss.detach(), generated=generated.detach())
