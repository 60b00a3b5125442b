"""Literal restatement of the one-step DDIM used by GenPercept.  TEST INFRASTRUCTURE ONLY.

Follows /root/reference/src/customized_modules/ddim.py:144-217 (betas / alphas_cumprod
construction, ``_get_variance``) and diffusers' ``DDIMScheduler.set_timesteps`` / ``step``
(v_prediction branch) with /root/reference/hf_configs/scheduler_beta_1.0_1.0/scheduler_config.json
(beta_start = beta_end = 1, scaled_linear, steps_offset 1, leading spacing, set_alpha_to_one False,
clip_sample False).  tests/test_oracle.py uses it to pin the collapse
``pred_original_sample == -model_output`` and ``timesteps == [1]``.
"""
import numpy as np
import torch


class DDIMOneStep:
    def __init__(self, num_train_timesteps=1000, beta_start=1.0, beta_end=1.0, steps_offset=1,
                 set_alpha_to_one=False):
        self.num_train_timesteps = num_train_timesteps
        self.steps_offset = steps_offset
        self.betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps,
                                    dtype=torch.float32) ** 2
        self.alphas = 1.0 - self.betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
        self.final_alpha_cumprod = torch.tensor(1.0) if set_alpha_to_one else self.alphas_cumprod[0]

    def set_timesteps(self, num_inference_steps):
        # "leading" spacing
        step_ratio = self.num_train_timesteps // num_inference_steps
        ts = (np.arange(0, num_inference_steps) * step_ratio).round()[::-1].copy().astype(np.int64)
        ts += self.steps_offset
        self.num_inference_steps = num_inference_steps
        self.timesteps = torch.from_numpy(ts)
        return self.timesteps

    def step(self, model_output, timestep, sample):
        prev_timestep = int(timestep) - self.num_train_timesteps // self.num_inference_steps
        alpha_prod_t = self.alphas_cumprod[int(timestep)]
        alpha_prod_t_prev = self.alphas_cumprod[prev_timestep] if prev_timestep >= 0 else self.final_alpha_cumprod
        beta_prod_t = 1 - alpha_prod_t
        # v_prediction
        pred_original_sample = (alpha_prod_t ** 0.5) * sample - (beta_prod_t ** 0.5) * model_output
        pred_epsilon = (alpha_prod_t ** 0.5) * model_output + (beta_prod_t ** 0.5) * sample
        # eta = 0 -> std_dev_t = 0 (variance from ddim.py:206-217 is multiplied by eta)
        pred_sample_direction = (1 - alpha_prod_t_prev) ** 0.5 * pred_epsilon
        prev_sample = alpha_prod_t_prev ** 0.5 * pred_original_sample + pred_sample_direction
        return prev_sample, pred_original_sample
