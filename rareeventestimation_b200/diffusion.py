"""The reference's diffusion problem (``problem/diffusion.py:245-252``): d = 150 KL modes, 512 elements."""
from .lsf import DiffusionLSF
from .problem import NormalProblem

parameter_dim = 150
lsf = DiffusionLSF(dim=parameter_dim)
diffusion_problem = NormalProblem(lsf, parameter_dim, 1, name=f"Diffusion Problem (d={parameter_dim})",
                                  prob_fail_true=1.682 * 1e-4)
