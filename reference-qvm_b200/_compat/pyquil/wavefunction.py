"""``Wavefunction`` result container (pyquil 1.x look-alike)."""
import numpy as np


class Wavefunction(object):
    def __init__(self, amplitude_vector):
        self.amplitudes = amplitude_vector

    def __len__(self):
        n = len(self.amplitudes)
        return int(n).bit_length() - 1

    def __iter__(self):
        return iter(self.amplitudes)

    def __getitem__(self, index):
        return self.amplitudes[index]

    def probabilities(self):
        a = np.asarray(self.amplitudes)
        return (a.real ** 2 + a.imag ** 2)

    def get_outcome_probs(self):
        n = len(self)
        return {format(i, "0%db" % max(n, 1)): float(p) for i, p in enumerate(self.probabilities())}
