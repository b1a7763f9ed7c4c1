"""Sweep: same interface as the reference (sweep.py:15-148), but every grid point is solved
in one batched GPU call instead of a Python loop over ``Circuit.diag``."""
from itertools import product

import numpy as np

DEC_TYPES = ['capacitive', 'inductive', 'quasiparticle', 'charge', 'cc', 'flux']


class Sweep:
    def __init__(self, cr, numEig, properties=None):
        self.cr = cr
        self.numEig = numEig
        self.properties = properties if properties is not None else ['efreq']
        self.type = None
        self.params = None
        self.grid = None
        self.efreq = None
        self.dec = {k: None for k in DEC_TYPES} if 'loss' in self.properties else None
        self.result = None

    def _run(self, kind, params, grid):
        dims = tuple(len(g) for g in grid)
        pts = list(product(*[range(d) for d in dims]))
        cols = [np.array([grid[i][idx[i]] for idx in pts], dtype=float) for i in range(len(grid))]
        if kind == 'flux':
            res = self.cr.sweep_diag(self.numEig, loop_fluxes={params[i]: cols[i] for i in range(len(grid))})
        else:
            res = self.cr.sweep_diag(self.numEig, charge_offsets={params[i]: cols[i] for i in range(len(grid))})
        self.result = res
        ef = res.efreqs.cpu().numpy()                      # [B][k]
        self.efreq = ef.T.reshape((self.numEig,) + dims)
        if 'loss' in self.properties:
            self.dec = {}
            for dt in DEC_TYPES:
                self.dec[dt] = np.asarray(res.dec_rate(dt, states=(1, 0))).reshape(dims)
        # leave the circuit at the last grid point, like the reference loop does
        for i, prm in enumerate(params):
            if kind == 'flux':
                prm.set_flux(grid[i][-1])
            else:
                self.cr.set_charge_offset(mode=prm, ng=grid[i][-1])
        return self.efreq, self.dec

    def sweepFlux(self, loops, grid, toFile=None, plotF=False):
        self.type, self.params, self.grid = 'sweepFlux', loops, grid
        out = self._run('flux', loops, grid)
        if toFile:
            self.save(toFile)
            return None
        return out

    def sweepCharge(self, modes, grid, toFile=None, plotF=False):
        self.type, self.params, self.grid = 'sweepCharge', modes, grid
        out = self._run('charge', modes, grid)
        if toFile:
            self.save(toFile)
            return None
        return out

    def save(self, toFile):
        import pickle
        with open(toFile, 'wb') as f:
            pickle.dump({'type': self.type, 'grid': self.grid, 'efreq': self.efreq, 'dec': self.dec}, f)
