"""Circuit builders shared by the tests (product package and oracle side by side)."""
import numpy as np

import sqcircuit_b200 as sq
from oracle import sqc_oracle as o


def zero_pi(kernels=None):
    lp = sq.Loop()
    C, CJ = sq.Capacitor(0.15, 'GHz'), sq.Capacitor(10, 'GHz')
    JJ, L = sq.Junction(5, 'GHz', loops=[lp]), sq.Inductor(0.13, 'GHz', loops=[lp])
    cr = sq.Circuit({(0, 1): [CJ, JJ], (0, 2): [L], (0, 3): [C], (1, 2): [C], (1, 3): [L],
                     (2, 3): [CJ, JJ]}, kernels=kernels)
    ol = o.OLoop()
    return cr, lp, o.zero_pi(ol), ol


def fluxonium(kernels=None):
    lp = sq.Loop()
    cap = sq.Capacitor(3.6, 'GHz', Q=1e6)
    ind = sq.Inductor(0.46, 'GHz', Q=500e6, loops=[lp])
    junc = sq.Junction(10.2, 'GHz', cap=cap, A=1e-7, x=3e-06, loops=[lp])
    cr = sq.Circuit({(0, 1): [ind, junc]}, flux_dist='all', kernels=kernels)
    ol = o.OLoop()
    return cr, lp, o.fluxonium(ol), ol


def transmon(kernels=None):
    cr = sq.Circuit({(0, 1): [sq.Capacitor(20, 'GHz'), sq.Junction(15.0, 'GHz', A=1e-7)]},
                    kernels=kernels)
    return cr, o.transmon()


def transmon_single(kernels=None):
    cr = sq.Circuit({(0, 1): [sq.Capacitor(7.746, 'fF', Q=1e6), sq.Junction(5, 'GHz')]}, kernels=kernels)
    oc = o.OCircuit({(0, 1): [o.OCap(7.746, 'fF', Q=1e6), o.OJJ(5, 'GHz')]})
    return cr, oc


def coupled_fluxonium_transmon(kernels=None):
    lp = sq.Loop()
    C = sq.Capacitor(1, 'GHz')
    cr = sq.Circuit({(0, 1): [sq.Inductor(1, 'GHz', loops=[lp]), sq.Junction(1, 'GHz', loops=[lp]), C],
                     (0, 2): [sq.Junction(1, 'GHz'), C], (1, 2): [sq.Capacitor(1, 'GHz')]}, kernels=kernels)
    ol = o.OLoop()
    Co = o.OCap(1, 'GHz')
    oc = o.OCircuit({(0, 1): [o.OInd(1, 'GHz', loops=[ol]), o.OJJ(1, 'GHz', loops=[ol]), Co],
                     (0, 2): [o.OJJ(1, 'GHz'), Co], (1, 2): [o.OCap(1, 'GHz')]})
    return cr, lp, oc, ol


def trt(kernels=None):
    cr = sq.Circuit({(0, 1): [sq.Capacitor(0.5, 'GHz'), sq.Inductor(8.0, 'GHz')],
                     (0, 2): [sq.Capacitor(0.30, 'GHz'), sq.Junction(14.0, 'GHz')],
                     (0, 3): [sq.Capacitor(0.27, 'GHz'), sq.Junction(17.0, 'GHz')],
                     (1, 2): [sq.Capacitor(12.0, 'GHz')], (1, 3): [sq.Capacitor(15.0, 'GHz')]},
                    kernels=kernels)
    oc = o.OCircuit({(0, 1): [o.OCap(0.5, 'GHz'), o.OInd(8.0, 'GHz')],
                     (0, 2): [o.OCap(0.30, 'GHz'), o.OJJ(14.0, 'GHz')],
                     (0, 3): [o.OCap(0.27, 'GHz'), o.OJJ(17.0, 'GHz')],
                     (1, 2): [o.OCap(12.0, 'GHz')], (1, 3): [o.OCap(15.0, 'GHz')]})
    return cr, oc


def trt_squid(kernels=None):
    """Transmon - resonator - flux-tunable transmon (the second transmon is a SQUID: two junctions in a loop):
    BASELINE configs[3] with a gate-charge AND a flux axis."""
    lp = sq.Loop()
    cr = sq.Circuit({(0, 1): [sq.Capacitor(0.5, 'GHz'), sq.Inductor(8.0, 'GHz')],
                     (0, 2): [sq.Capacitor(0.30, 'GHz'), sq.Junction(14.0, 'GHz')],
                     (0, 3): [sq.Capacitor(0.27, 'GHz'), sq.Junction(9.0, 'GHz', loops=[lp]),
                              sq.Junction(8.0, 'GHz', loops=[lp])],
                     (1, 2): [sq.Capacitor(12.0, 'GHz')], (1, 3): [sq.Capacitor(15.0, 'GHz')]}, kernels=kernels)
    ol = o.OLoop()
    oc = o.OCircuit({(0, 1): [o.OCap(0.5, 'GHz'), o.OInd(8.0, 'GHz')],
                     (0, 2): [o.OCap(0.30, 'GHz'), o.OJJ(14.0, 'GHz')],
                     (0, 3): [o.OCap(0.27, 'GHz'), o.OJJ(9.0, 'GHz', loops=[ol]), o.OJJ(8.0, 'GHz', loops=[ol])],
                     (1, 2): [o.OCap(12.0, 'GHz')], (1, 3): [o.OCap(15.0, 'GHz')]})
    return cr, lp, oc, ol


def inductively_shunted(kernels=None):
    lp = sq.Loop()
    cr = sq.Circuit({(0, 1): [sq.Capacitor(20.3, 'fF')], (1, 2): [sq.Inductor(15.6, 'nH')],
                     (0, 2): [sq.Inductor(4.5, 'nH', loops=[lp])], (2, 3): [sq.Inductor(386, 'nH', loops=[lp])],
                     (0, 3): [sq.Junction(6.2, 'GHz', loops=[lp]), sq.Capacitor(5.3, 'fF')]}, kernels=kernels)
    ol = o.OLoop()
    oc = o.OCircuit({(0, 1): [o.OCap(20.3, 'fF')], (1, 2): [o.OInd(15.6, 'nH')],
                     (0, 2): [o.OInd(4.5, 'nH', loops=[ol])], (2, 3): [o.OInd(386, 'nH', loops=[ol])],
                     (0, 3): [o.OJJ(6.2, 'GHz', loops=[ol]), o.OCap(5.3, 'fF')]})
    return cr, lp, oc, ol
