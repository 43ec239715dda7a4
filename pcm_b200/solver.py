"""Drop-in mirrors of the reference's solver-side Python API (same names, argument meaning and
error behaviour), backed by the fused CUDA kernels:

  DDIMSolver                                   train_pcm_lora_sd15.py:289-341
  predicted_origin / extract_into_tensor       :268-286
  append_dims                                  :240-247
  scalings_for_boundary_conditions_{target,online}   :250-259
  PCMNoiseScheduler.add_noise / noise_travel   scheduling_ddpm_modified.py:500-554

Tensor arguments must be CUDA tensors (the product path has no CPU fallback); results have the
reference's dtypes (solver steps return float64 because the reference's alpha table is float64).
"""
import numpy as np
import torch

from . import ops


def append_dims(x, target_dims):
    """Appends dimensions to the end of a tensor until it has target_dims dimensions."""
    dims_to_append = target_dims - x.ndim
    if dims_to_append < 0:
        raise ValueError(f"input has {x.ndim} dims but target_dims is {target_dims}, which is less")
    return x[(...,) + (None,) * dims_to_append]


def scalings_for_boundary_conditions_target(index, selected_indices):
    c_skip = torch.isin(index, selected_indices).float()
    return c_skip, 1.0 - c_skip


def scalings_for_boundary_conditions_online(index, selected_indices):
    return torch.zeros_like(index).float(), torch.ones_like(index).float()


def extract_into_tensor(a, t, x_shape):
    b = t.shape[0]
    return a.gather(-1, t).reshape(b, *((1,) * (len(x_shape) - 1)))


def _require_cuda(*ts):
    for t in ts:
        if not t.is_cuda:
            raise RuntimeError("pcm_b200 solver ops need CUDA tensors (no CPU fallback)")


def _axpby64(x, y, ca, cb):
    """ca[b]*x + cb[b]*y -> float64, one fused kernel (pcm_axpby_f64)."""
    _require_cuda(x, y)
    B = x.shape[0]
    xf, yf = x.float().contiguous(), y.float().contiguous()
    out = torch.empty(x.shape, device=x.device, dtype=torch.float64)
    ops._call("pcm_axpby_f64", xf.data_ptr(), yf.data_ptr(), ca.contiguous().data_ptr(),
              cb.contiguous().data_ptr(), xf.numel() // B, B, out.data_ptr())
    return out


def predicted_origin(model_output, timesteps, sample, prediction_type, alphas, sigmas):
    if prediction_type not in ("epsilon", "v_prediction"):
        raise ValueError(f"Prediction type {prediction_type} currently not supported.")
    s = sigmas.gather(-1, timesteps).double()
    a = alphas.gather(-1, timesteps).double()
    if prediction_type == "epsilon":  # (sample - sigma * eps) / alpha
        return _axpby64(sample, model_output, 1.0 / a, -s / a).float()
    return _axpby64(sample, model_output, a, -s).float()  # alpha * sample - sigma * v


class DDIMSolver:
    def __init__(self, alpha_cumprods, timesteps=1000, ddim_timesteps=50):
        self.step_ratio = timesteps // ddim_timesteps
        ts = (np.arange(1, ddim_timesteps + 1) * self.step_ratio).round().astype(np.int64) - 1
        prev_ts = np.concatenate([[0], ts[:-1]]).astype(np.int64)
        # the reference builds the "prev" alpha table from a Python list -> float64
        prev_alpha = np.asarray([float(alpha_cumprods[0])] + [float(v) for v in alpha_cumprods[ts[:-1]]])
        self.ddim_timesteps = torch.from_numpy(ts).long()
        self.ddim_timesteps_prev = torch.from_numpy(prev_ts).long()
        self.ddim_alpha_cumprods = torch.from_numpy(np.asarray(alpha_cumprods[ts]))
        self.ddim_alpha_cumprods_prev = torch.from_numpy(prev_alpha)

    def to(self, device):
        self.ddim_timesteps = self.ddim_timesteps.to(device)
        self.ddim_timesteps_prev = self.ddim_timesteps_prev.to(device)
        self.ddim_alpha_cumprods = self.ddim_alpha_cumprods.to(device)
        self.ddim_alpha_cumprods_prev = self.ddim_alpha_cumprods_prev.to(device)
        return self

    def _jump(self, pred_x0, pred_noise, index):
        a = self.ddim_alpha_cumprods_prev.gather(-1, index)
        return _axpby64(pred_x0, pred_noise, a.sqrt(), (1.0 - a).sqrt())

    def ddim_step(self, pred_x0, pred_noise, timestep_index):
        return self._jump(pred_x0, pred_noise, timestep_index)

    def phase_start_index(self, timestep_index, multiphase):
        inf = np.floor(np.linspace(0, len(self.ddim_timesteps), num=multiphase, endpoint=False)).astype(np.int64)
        inf = torch.from_numpy(inf).long().to(self.ddim_timesteps.device)
        pos = (timestep_index[:, None] >= inf[None, :]).sum(1) - 1
        return inf[pos]

    def ddim_style_multiphase_pred(self, pred_x0, pred_noise, timestep_index, multiphase):
        p = self.phase_start_index(timestep_index, multiphase)
        return self._jump(pred_x0, pred_noise, p), self.ddim_timesteps_prev[p]

    # SDXL script spelling (train_pcm_lora_sdxl_adv.py:345)
    ddim_style_multiphase = ddim_style_multiphase_pred


class _Cfg:
    def __init__(self, **kw):
        self.__dict__.update(kw)


class PCMNoiseScheduler:
    """The slice of (modified) DDPMScheduler the training loop touches: `alphas_cumprod`,
    `config.{num_train_timesteps,prediction_type}`, `add_noise`, `noise_travel`."""

    def __init__(self, num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012,
                 beta_schedule="scaled_linear", prediction_type="epsilon"):
        if beta_schedule == "scaled_linear":
            betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float32) ** 2
        elif beta_schedule == "linear":
            betas = torch.linspace(beta_start, beta_end, num_train_timesteps, dtype=torch.float32)
        else:
            raise NotImplementedError(f"{beta_schedule} does is not implemented for {self.__class__}")
        self.betas = betas
        self.alphas = 1.0 - betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
        self.config = _Cfg(num_train_timesteps=num_train_timesteps, prediction_type=prediction_type,
                           beta_start=beta_start, beta_end=beta_end, beta_schedule=beta_schedule)

    def add_noise(self, original_samples, noise, timesteps):
        _require_cuda(original_samples, noise)
        self.alphas_cumprod = self.alphas_cumprod.to(device=original_samples.device)
        ac = self.alphas_cumprod.to(dtype=original_samples.dtype)[timesteps.to(original_samples.device)]
        out = _axpby64(original_samples, noise, (ac ** 0.5).double(), ((1 - ac) ** 0.5).double())
        return out.to(original_samples.dtype)

    def noise_travel(self, current_samples, noise, current_timesteps, target_timesteps):
        _require_cuda(current_samples, noise)
        x = current_samples.float().contiguous()
        n = noise.float().contiguous()
        acp = self.alphas_cumprod.to(device=x.device, dtype=torch.float32).contiguous()
        out = torch.empty_like(x)
        B = x.shape[0]
        ops._call("pcm_noise_travel", x.data_ptr(), n.data_ptr(), acp.data_ptr(),
                  current_timesteps.to(x.device).long().contiguous().data_ptr(),
                  target_timesteps.to(x.device).long().contiguous().data_ptr(), x.numel() // B, B, out.data_ptr())
        return out.to(current_samples.dtype)


def sample_adv_timesteps(end_timesteps, num_train_timesteps, multiphase, generator=None):
    """Per-sample adversarial timesteps of the adversarial PCM variant
    (train_pcm_lora_sd15_adv.py:1288-1298): adv_t[i] ~ U{end_t[i], ..., end_t[i] + T // multiphase - 1}.
    The reference draws them in a Python loop with one `.item()` host sync per sample; this is the same
    distribution drawn on the device in one call (no sync), ready for `PCMNoiseScheduler.noise_travel`."""
    span = num_train_timesteps // multiphase
    off = torch.randint(0, span, end_timesteps.shape, device=end_timesteps.device, dtype=end_timesteps.dtype,
                        generator=generator)
    return end_timesteps + off
