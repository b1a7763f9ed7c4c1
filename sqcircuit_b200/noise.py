"""Environment parameters of the decoherence models (reference noise.py)."""
import numpy as np

from . import units as unt

ENV = {
    'T': 0.015,
    'omega_low': 2 * np.pi,
    'omega_high': 2 * np.pi * 3 * 1e9,
    't_exp': 10e-6,
}


def set_temp(T):
    ENV['T'] = T


def set_low_freq(value, unit):
    ENV['omega_low'] = 2 * np.pi * value * unt.freq_list[unit]


def set_high_freq(value, unit):
    ENV['omega_high'] = 2 * np.pi * value * unt.freq_list[unit]


def set_t_exp(value, unit):
    ENV['t_exp'] = value * unt.time_list[unit]


def reset_to_default():
    set_temp(0.015)
    set_low_freq(1, 'Hz')
    set_high_freq(3, 'GHz')
    set_t_exp(10, 'us')
