"""ertdiff_b200 -- B200-native ensemble posterior sampling for the ERT conditional diffusion
model (pnnl/ERT-Conditional-Diffusion-Model).  Drop-in host API over ``libertdiff_b200.so``:

    model = ConditionalDiffusionModel(param_dim=29, hidden_dim=128).to("cuda")
    model.load_state_dict(reference_state_dict)
    betas, alphas, alpha_bar = get_diffusion_schedule(T)
    x0 = sample_model(model, condition, T, betas, alphas, alpha_bar, 29, "cuda")
    stats = ensemble_statistics(x0)

No CPU fallback: importing works anywhere the shared library exists, computing needs a GPU.
"""
from . import _lib
from ._lib import ErtdiffError, LIB_PATH
from .model import ConditionalDiffusionModel
from .sampler import (get_diffusion_schedule, get_timestep_embedding, sample_model,
                      sample_ensemble, run_chain, step_coefficients, posterior_update,
                      philox_normal, debug_umma_gemm)
from .stats import (ensemble_moments, ensemble_mean, ensemble_std, ensemble_var,
                    ensemble_percentile, ensemble_kde_mode, ensemble_statistics, ensemble_summary, ensemble_summary_packed, uq_calibration, misfit_metrics,
                    wasserstein_distance)
from .transforms import untransform_and_check, inverse_transform, check_param_bounds
from .checkpoint import load_best_model, save_checkpoint
from . import parallel


def launch_count(reset: bool = False) -> int:
    """Kernels launched by the library so far (bench.py's ``gpu_launches``)."""
    lib = _lib.load()
    n = int(lib.ertdiff_launch_count())
    if reset:
        lib.ertdiff_launch_count_reset()
    return n


__all__ = [n for n in dir() if not n.startswith("_")]
