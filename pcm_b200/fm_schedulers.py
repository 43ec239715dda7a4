"""Flow-matching (SD3) side of the reference's solver / scheduler API, same names, argument meaning,
state handling and error behaviour, backed by the fused CUDA kernels (no CPU fallback):

  EulerSolver                    train_pcm_lora_sd3.py:160-226   (training: sigma tables, euler_step,
                                                                  euler_style_multiphase_pred)
  PCMFMDeterministicScheduler    pcm_fm_deterministic_scheduler.py:35-242   (inference)
  PCMFMStochasticScheduler       pcm_fm_stochastic_scheduler.py:35-243      (inference, re-noising step)

These are the "PCMScheduler" objects of the reference (there is no class of that literal name,
SURVEY section 0.3).  The diffusers mixins (ConfigMixin / SchedulerMixin) are not available offline;
`.config` exposes the registered constructor arguments the same way.  The MMDiT network itself is out
of scope; these classes are the solver arithmetic around it.
"""
from dataclasses import dataclass
from types import SimpleNamespace

import numpy as np
import torch

from . import ops
from .solver import _axpby64, _require_cuda


class EulerSolver:
    def __init__(self, sigmas, timesteps=1000, euler_timesteps=50):
        self.step_ratio = timesteps // euler_timesteps
        self.euler_timesteps = (np.arange(1, euler_timesteps + 1) * self.step_ratio).round().astype(np.int64) - 1
        self.euler_timesteps_prev = np.asarray([0] + self.euler_timesteps[:-1].tolist())
        self.sigmas = sigmas[self.euler_timesteps]
        # built from a Python list like the reference: float64
        self.sigmas_prev = np.asarray([sigmas[0]] + sigmas[self.euler_timesteps[:-1]].tolist())
        self.euler_timesteps = torch.from_numpy(self.euler_timesteps).long()
        self.euler_timesteps_prev = torch.from_numpy(self.euler_timesteps_prev).long()
        self.sigmas = torch.from_numpy(self.sigmas)
        self.sigmas_prev = torch.from_numpy(self.sigmas_prev)

    def to(self, device):
        self.euler_timesteps = self.euler_timesteps.to(device)
        self.euler_timesteps_prev = self.euler_timesteps_prev.to(device)
        self.sigmas = self.sigmas.to(device)
        self.sigmas_prev = self.sigmas_prev.to(device)
        return self

    def _step(self, sample, model_pred, sigma, sigma_prev):
        # x + (sigma_prev - sigma) * v with the reference's type promotion (float64 result)
        coef = sigma_prev.double() - sigma.double()
        one = torch.ones_like(coef)
        return _axpby64(sample, model_pred, one, coef)

    def euler_step(self, sample, model_pred, timestep_index):
        _require_cuda(sample, model_pred)
        sigma = self.sigmas.gather(-1, timestep_index)
        sigma_prev = self.sigmas_prev.gather(-1, timestep_index)
        return self._step(sample, model_pred, sigma, sigma_prev)

    def euler_style_multiphase_pred(self, sample, model_pred, timestep_index, multiphase, is_target=False):
        _require_cuda(sample, model_pred)
        inference_indices = np.linspace(0, len(self.euler_timesteps), num=multiphase, endpoint=False)
        inference_indices = np.floor(inference_indices).astype(np.int64)
        inference_indices = torch.from_numpy(inference_indices).long().to(self.euler_timesteps.device)
        expanded = timestep_index.unsqueeze(1).expand(-1, inference_indices.size(0))
        valid = expanded >= inference_indices
        last_valid = valid.flip(dims=[1]).long().argmax(dim=1)
        last_valid = inference_indices.size(0) - 1 - last_valid
        timestep_index_end = inference_indices[last_valid]
        sigma = (self.sigmas_prev if is_target else self.sigmas).gather(-1, timestep_index)
        sigma_prev = self.sigmas_prev.gather(-1, timestep_index_end)
        return self._step(sample, model_pred, sigma, sigma_prev), timestep_index_end


@dataclass
class PCMFMSchedulerOutput:
    prev_sample: torch.Tensor


PCMFMDeterministicSchedulerOutput = PCMFMSchedulerOutput
PCMFMStochasticSchedulerOutput = PCMFMSchedulerOutput


class _PCMFMScheduler:
    order = 1
    _stochastic = False

    def __init__(self, num_train_timesteps: int = 1000, shift: float = 1.0, pcm_timesteps: int = 50):
        self.config = SimpleNamespace(num_train_timesteps=num_train_timesteps, shift=shift, pcm_timesteps=pcm_timesteps)
        timesteps = np.linspace(1, num_train_timesteps, num_train_timesteps, dtype=np.float32)[::-1].copy()
        timesteps = torch.from_numpy(timesteps).to(dtype=torch.float32)
        sigmas = timesteps / num_train_timesteps
        sigmas = shift * sigmas / (1 + (shift - 1) * sigmas)
        self.euler_timesteps = (np.arange(1, pcm_timesteps + 1) * (num_train_timesteps // pcm_timesteps)
                                ).round().astype(np.int64) - 1
        self.sigmas = sigmas.numpy()[::-1][self.euler_timesteps]
        self.sigmas = torch.from_numpy(self.sigmas[::-1].copy())
        self.timesteps = self.sigmas * num_train_timesteps
        self._step_index = None
        self._begin_index = None
        self.sigmas = self.sigmas.to("cpu")
        self.sigma_min = self.sigmas[-1].item()
        self.sigma_max = self.sigmas[0].item()

    @property
    def step_index(self):
        return self._step_index

    @property
    def begin_index(self):
        return self._begin_index

    def set_begin_index(self, begin_index: int = 0):
        self._begin_index = begin_index

    def _sigma_to_t(self, sigma):
        return sigma * self.config.num_train_timesteps

    def set_timesteps(self, num_inference_steps: int, device=None):
        self.num_inference_steps = num_inference_steps
        inference_indices = np.linspace(0, self.config.pcm_timesteps, num=num_inference_steps, endpoint=False)
        inference_indices = torch.from_numpy(np.floor(inference_indices).astype(np.int64)).long()
        self.sigmas_ = self.sigmas[inference_indices]
        timesteps = self.sigmas_ * self.config.num_train_timesteps
        self.timesteps = timesteps.to(device=device)
        self.sigmas_ = torch.cat([self.sigmas_, torch.zeros(1, device=self.sigmas_.device)])
        self._step_index = None
        self._begin_index = None

    def index_for_timestep(self, timestep, schedule_timesteps=None):
        if schedule_timesteps is None:
            schedule_timesteps = self.timesteps
        indices = (schedule_timesteps == timestep).nonzero()
        pos = 1 if len(indices) > 1 else 0
        return indices[pos].item()

    def _init_step_index(self, timestep):
        if self.begin_index is None:
            if isinstance(timestep, torch.Tensor):
                timestep = timestep.to(self.timesteps.device)
            self._step_index = self.index_for_timestep(timestep)
        else:
            self._step_index = self._begin_index

    @staticmethod
    def _launch(mode, x, v, z, sigma, sigma_next):
        _require_cuda(x)
        B = x.shape[0]
        xf = x.float().contiguous()
        vf = v.float().contiguous() if v is not None else xf
        zf = z.float().contiguous() if z is not None else xf
        s = torch.full((B,), float(sigma), device=x.device, dtype=torch.float32)
        sn = torch.full((B,), float(sigma_next), device=x.device, dtype=torch.float32)
        out = torch.empty_like(xf)
        ops._call("pcm_fm_step", xf.data_ptr(), vf.data_ptr(), zf.data_ptr(), s.data_ptr(), sn.data_ptr(),
                  xf.numel() // B, B, mode, out.data_ptr())
        return out

    def scale_noise(self, sample, timestep, noise=None):
        """Forward process in flow-matching: sigma * noise + (1 - sigma) * sample."""
        if self.step_index is None:
            self._init_step_index(timestep)
        sigma = self.sigmas[self.step_index]
        return self._launch(2, sample, None, noise, sigma, 0.0).to(torch.promote_types(sample.dtype, noise.dtype))

    def step(self, model_output, timestep, sample, generator=None, return_dict=True, noise=None):
        if isinstance(timestep, int) or isinstance(timestep, torch.IntTensor) or isinstance(timestep, torch.LongTensor):
            raise ValueError(
                "Passing integer indices (e.g. from `enumerate(timesteps)`) as timesteps to"
                " `EulerDiscreteScheduler.step()` is not supported. Make sure to pass"
                " one of the `scheduler.timesteps` as a timestep.")
        if self.step_index is None:
            self._init_step_index(timestep)
        sigma = self.sigmas_[self.step_index]
        sigma_next = self.sigmas_[self.step_index + 1]
        if self._stochastic:
            if noise is None:   # torch.randn_like(denoised) in the reference
                noise = torch.randn(sample.shape, device=sample.device, dtype=torch.float32, generator=generator)
            prev = self._launch(1, sample, model_output, noise, sigma, sigma_next)
        else:
            prev = self._launch(0, sample, model_output, None, sigma, sigma_next)
        prev = prev.to(model_output.dtype)
        self._step_index += 1
        if not return_dict:
            return (prev,)
        return PCMFMSchedulerOutput(prev_sample=prev)

    def __len__(self):
        return self.config.num_train_timesteps


class PCMFMDeterministicScheduler(_PCMFMScheduler):
    """Deterministic phased Euler sampler (pcm_fm_deterministic_scheduler.py:35-242)."""
    _stochastic = False


class PCMFMStochasticScheduler(_PCMFMScheduler):
    """Phase-boundary re-noising sampler (pcm_fm_stochastic_scheduler.py:35-243)."""
    _stochastic = True
