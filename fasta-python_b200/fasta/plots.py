"""Optional matplotlib helpers with the reference's names (reference fasta/plots.py:13-86).

Visualisation is outside the hot path; matplotlib is imported lazily so that ``import fasta``
never depends on it (the reference's unconditional import is what breaks it where matplotlib is
absent, reference fasta/__init__.py:26).
"""

import numpy as np


def _plt():
    from matplotlib import pyplot as plt
    return plt


def plot_convergence(title, solutions, labels):
    plt = _plt()
    figure, axes = plt.subplots(1, 2)
    figure.suptitle(title)
    for sol, label in zip(solutions, labels):
        n = sol.iteration_count
        axes[0].semilogy(np.arange(n), sol.residuals[:n], label=label)
        if sol.objectives is not None:
            axes[1].semilogy(np.arange(n + 1), sol.objectives[:n + 1], label=label)
    axes[0].set_xlabel("Iteration #"), axes[0].set_ylabel("Residual"), axes[0].legend()
    axes[1].set_xlabel("Iteration #"), axes[1].set_ylabel("Objective"), axes[1].legend()
    return figure


def plot_signals(title, original, recovered):
    plt = _plt()
    figure, axes = plt.subplots(2, 1, sharex=True)
    figure.suptitle(title)
    axes[0].stem(np.asarray(original)), axes[0].set_title("Original")
    axes[1].stem(np.asarray(recovered)), axes[1].set_title("Recovered")
    return figure


def plot_matrices(title, original, recovered):
    plt = _plt()
    figure, axes = plt.subplots(1, 2)
    figure.suptitle(title)
    axes[0].imshow(np.asarray(original), cmap="gray"), axes[0].set_title("Original")
    axes[1].imshow(np.asarray(recovered), cmap="gray"), axes[1].set_title("Recovered")
    return figure
