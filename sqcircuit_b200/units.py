"""Physical constants and unit tables (values as in reference units.py:3-29)."""
from .exceptions import raise_unit_error

henry_list = {'H': 1.0, 'mH': 1.0e-3, 'uH': 1.0e-6, 'nH': 1.0e-9, 'pH': 1.0e-12, 'fH': 1.0e-15}
farad_list = {'F': 1, 'mF': 1.0e-3, 'uF': 1.0e-6, 'nF': 1.0e-9, 'pF': 1.0e-12, 'fF': 1.0e-15}
freq_list = {'Hz': 1.0, 'kHz': 1.0e3, 'MHz': 1.0e6, 'GHz': 1.0e9, 'THz': 1.0e12}
time_list = {'s': 1.0, 'ms': 1.0e-3, 'us': 1.0e-6, 'ns': 1.0e-9, 'ps': 1.0e-12, 'fs': 1.0e-15}

hbar = 1.0545718e-34
Phi0 = 2.067833e-15
e = 1.6021766e-19
k_B = 1.38e-23

_state = {'freq': freq_list['GHz'], 'cap': 'GHz', 'ind': 'GHz', 'JJ': 'GHz'}


def set_unit_freq(unit):
    if unit not in freq_list:
        raise_unit_error()
    _state['freq'] = freq_list[unit]


def get_unit_freq():
    return _state['freq']


def set_unit_cap(unit):
    if unit not in freq_list and unit not in farad_list:
        raise_unit_error()
    _state['cap'] = unit


def get_unit_cap():
    return _state['cap']


def set_unit_ind(unit):
    if unit not in freq_list and unit not in henry_list:
        raise_unit_error()
    _state['ind'] = unit


def get_unit_ind():
    return _state['ind']


def set_unit_JJ(unit):
    if unit not in freq_list:
        raise_unit_error()
    _state['JJ'] = unit


def get_unit_JJ():
    return _state['JJ']
