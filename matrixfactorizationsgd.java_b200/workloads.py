"""The five BASELINE.json configs as data (shapes from BASELINE.json `configs`; lr/lambda/epochs and
the skew parameters are the pins of SURVEY.md section 8d)."""
from collections import namedtuple

Workload = namedtuple("Workload", "name n_users n_items n_ratings k epochs lr lambda_ log2_alpha_user c_user "
                                  "log2_alpha_item c_item gpus")

SEED = 20261018

WORKLOADS = {
    "ml100k": Workload("ml100k-shaped", 943, 1682, 100_000, 32, 20, 0.01, 0.05, 2, 0.25, 3, 0.375, (1,)),
    "ml20m": Workload("ml20m-shaped", 138_000, 27_000, 20_000_000, 128, 10, 0.005, 0.05, 2, 0.25, 3, 0.375, (1,)),
    "netflix": Workload("netflix-shaped", 480_000, 17_800, 100_000_000, 128, 10, 0.005, 0.05, 2, 0.25, 3, 0.375,
                        (1, 2, 4, 8)),
    "yahoo": Workload("yahoo-r2-shaped", 1_800_000, 136_000, 700_000_000, 128, 5, 0.005, 0.05, 2, 0.25, 3, 0.375, (8,)),
    "powerlaw": Workload("power-law-heavy", 10_000_000, 1_000_000, 2_000_000_000, 64, 3, 0.005, 0.05, 2, 0.25, 4, 0.375,
                         (8,)),
}


def bytes_per_update(k):
    """Algorithmic bytes of one SGD update: 12-B record + read and write of p_u and q_i."""
    return 12 + 16 * k
