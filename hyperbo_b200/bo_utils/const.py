"""Name -> callable registries -- mirrors hyperbo/bo_utils/const.py:22-50."""
from hyperbo_b200.bo_utils import acfun
from hyperbo_b200.bo_utils import data
from hyperbo_b200.gp_utils import kernel
from hyperbo_b200.gp_utils import mean

MEAN = {
    "constant": mean.constant,
    "linear": mean.linear,
    "linear_mlp": mean.linear_mlp,
    "zero": mean.zero,
}

KERNEL = {
    "squared_exponential": kernel.squared_exponential,
    "matern32": kernel.matern32,
    "matern52": kernel.matern52,
    "dot_product": kernel.dot_product,
    "dot_product_mlp": kernel.dot_product_mlp,
}

ACFUN = {
    "expected_improvement": acfun.expected_improvement,
    "probability_of_improvement": acfun.probability_of_improvement,
    "ucb3": acfun.ucb3,
    "random_search": acfun.random_search,
    "ucb2": acfun.ucb2,
    "ucb": acfun.ucb,
}

ACFUN_SUB = {
    "expected_improvement": acfun.expected_improvement_sub,
    "probability_of_improvement": acfun.probability_of_improvement_sub,
    "ucb": acfun.ucb_sub,
}

EPS = 1e-6

HYPERBO_DATASETS = {"random": data.random}  # const.py:54-59 (pd1 needs files)

# method-name constants (const.py:63-81)
RAND = "rand"
STBO = "stbo"
MTBO = "mtbo"
STBOV = "gp"
HBO = "hyperbo"
HBO_SS = "hyperbo_ss"
HBO_NLL = "hyperbo_nll"
HBO_NLLKL = "hyperbo_nllkl"
HBO_NLLEUC = "hyperbo_nlleuc"
USE_HGP = [HBO_SS]
