"""The caller of the hot path in training: one relativistic-average GAN step, restating
``SRRaGANModel.optimize_parameters`` (codes/models/SRRaGAN_model.py:113-186) for the ESRGAN+ recipe without the
perceptual branch (gan_type 'vanilla' = BCEWithLogits, models/modules/loss.py:5-38; L1 pixel loss :31-39; two Adam
optimisers :82-89).  The reference's own solver runs unchanged on the drop-in classes; this
restatement exists so that bench.py and the tests can drive the step on a box that has no copy of the reference.

On CUDA the solver arithmetic is native too (esrganplus_b200/solver.py): both Adam updates
are one kernel each over flat parameter / gradient / moment buffers (``FlatAdam``, a ``torch.optim.Optimizer``, so the
reference's ``MultiStepLR`` drives it), the relativistic BCE terms and the L1 loss are single launches with analytic
gradients.  With ``netF`` (a ``VGGFeatureExtractor``) the step also carries the perceptual term
of the shipped recipe (``feature_weight`` 1, L1; SRRaGAN_model.py:127-131).  ``native_solver=False`` keeps torch's own Adam / BCE / L1 (what the reference runs).  With ``data_parallel``
modules the two backward passes all-reduce their flat gradient buffers — nothing else is exchanged.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.nn.functional as F


class GanTrainStep:
    def __init__(self, netG, netD, lr_G: float = 1e-4, lr_D: float = 1e-4, beta1_G: float = 0.9, beta1_D: float = 0.9,
                 pixel_weight: float = 1e-2, gan_weight: float = 5e-3, weight_decay_G: float = 0.0,
                 weight_decay_D: float = 0.0, native_solver: Optional[bool] = None, lr_steps=None, lr_gamma: float = 0.5,
                 netF=None, feature_weight: float = 1.0):
        self.netG, self.netD, self.netF = netG, netD, netF
        self.l_pix_w, self.l_gan_w, self.l_fea_w = pixel_weight, gan_weight, feature_weight
        on_cuda = all(p.is_cuda for p in netG.parameters())
        self.native = on_cuda if native_solver is None else bool(native_solver)
        pg = [p for p in netG.parameters() if p.requires_grad]
        if self.native:
            from .autograd import enable_flat_grads
            from .solver import FlatAdam
            enable_flat_grads(netG)   # the backward points p.grad at slices of one persistent flat buffer (autograd.py)
            # SRRaGAN_model.py:77-89 as one kernel per network over flat storage
            self.optimizer_G = FlatAdam(pg, lr=lr_G, weight_decay=weight_decay_G, betas=(beta1_G, 0.999), module=netG)
            self.optimizer_D = FlatAdam(list(netD.parameters()), lr=lr_D, weight_decay=weight_decay_D, betas=(beta1_D, 0.999),
                                        module=netD)
        else:
            # fused=True is the same Adam arithmetic as one multi-tensor library kernel (771 + 69 tensors)
            self.optimizer_G = torch.optim.Adam(pg, lr=lr_G, weight_decay=weight_decay_G, betas=(beta1_G, 0.999), fused=on_cuda)
            self.optimizer_D = torch.optim.Adam(netD.parameters(), lr=lr_D, weight_decay=weight_decay_D, betas=(beta1_D, 0.999),
                                                fused=on_cuda)
        # lr_scheme MultiStepLR (SRRaGAN_model.py:91-95; train_ESRGANplus.json: lr_steps [50000, 100000, 200000, 300000], gamma 0.5)
        self.schedulers = []
        if lr_steps:
            for opt in (self.optimizer_G, self.optimizer_D):
                self.schedulers.append(torch.optim.lr_scheduler.MultiStepLR(opt, list(lr_steps), lr_gamma))
        self.log: Dict[str, torch.Tensor] = {}
        self.fake_H: Optional[torch.Tensor] = None

    @staticmethod
    def _gan(pred: torch.Tensor, target_is_real: bool) -> torch.Tensor:
        """GANLoss('vanilla', 1.0, 0.0) (loss.py:11-38): BCEWithLogits against a constant label."""
        return F.binary_cross_entropy_with_logits(pred, torch.full_like(pred, 1.0 if target_is_real else 0.0))

    def step(self, var_L: torch.Tensor, var_H: torch.Tensor, var_ref: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
        var_ref = var_H if var_ref is None else var_ref
        netG, netD = self.netG, self.netD
        # ---- G (SRRaGAN_model.py:115-141): D frozen, its data gradient flows into G
        for p in netD.parameters():
            p.requires_grad = False
        self.optimizer_G.zero_grad()
        self.fake_H = netG(var_L)
        pred_g_fake = netD(self.fake_H)
        pred_d_real = netD(var_ref).detach()
        if self.native:
            from .solver import l1_loss, ragan_bce_terms
            l_g_pix = self.l_pix_w * l1_loss(self.fake_H, var_H)
            a, b = ragan_bce_terms(pred_d_real, pred_g_fake, 0.0, 1.0)
            l_g_gan = self.l_gan_w * (a + b) / 2
        else:
            l_g_pix = self.l_pix_w * F.l1_loss(self.fake_H, var_H)
            l_g_gan = self.l_gan_w * (self._gan(pred_d_real - torch.mean(pred_g_fake), False) +
                                      self._gan(pred_g_fake - torch.mean(pred_d_real), True)) / 2
        l_g_total = l_g_pix + l_g_gan
        l_g_fea = None
        if self.netF is not None:                                # SRRaGAN_model.py:127-131
            real_fea = self.netF(var_H).detach()
            fake_fea = self.netF(self.fake_H)
            if self.native:
                l_g_fea = self.l_fea_w * l1_loss(fake_fea, real_fea)
            else:
                l_g_fea = self.l_fea_w * F.l1_loss(fake_fea, real_fea)
            l_g_total = l_g_total + l_g_fea
        l_g_total.backward()
        self.optimizer_G.step()
        # ---- D (SRRaGAN_model.py:143-168)
        for p in netD.parameters():
            p.requires_grad = True
        self.optimizer_D.zero_grad()
        pred_d_real = netD(var_ref)
        pred_d_fake = netD(self.fake_H.detach())
        if self.native:
            l_d_real, l_d_fake = ragan_bce_terms(pred_d_real, pred_d_fake, 1.0, 0.0)
        else:
            l_d_real = self._gan(pred_d_real - torch.mean(pred_d_fake), True)
            l_d_fake = self._gan(pred_d_fake - torch.mean(pred_d_real), False)
        ((l_d_real + l_d_fake) / 2).backward()
        self.optimizer_D.step()
        for sch in self.schedulers:   # base_model.update_learning_rate(), called once per iteration by train.py
            sch.step()
        # the reference calls .item() on each of these (six host syncs per step, :171-186); kept as device scalars here
        self.log = {"l_g_pix": l_g_pix.detach(), "l_g_gan": l_g_gan.detach(), "l_d_real": l_d_real.detach(),
                    "l_d_fake": l_d_fake.detach(), "D_real": pred_d_real.detach().mean(), "D_fake": pred_d_fake.detach().mean()}
        if l_g_fea is not None:
            self.log["l_g_fea"] = l_g_fea.detach()
        return self.log
