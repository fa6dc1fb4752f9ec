"""Trajectory file reader -- mirror of ``ves/utils.py`` (``TrajectoryReader``, ``:12-69``).

Reads the text files ``SingleParticleSimulation`` writes (``# t[ps] x y z`` / ``# t PE KE``).  Unlike
upstream, ``skip`` really skips frames (upstream never increments its counter) and
matplotlib is not imported.
"""
import numpy as np


class TrajectoryReader:
    def __init__(self, traj_file, comment_char='#', format='txyz'):
        self.traj_file = traj_file
        self.comment_char = comment_char
        self.format = format

    def _rows(self, skip=1):
        rows = []
        count = 0
        with open(self.traj_file) as fh:
            for line in fh:
                if not line.strip() or line.lstrip().startswith(self.comment_char):
                    continue
                if count % skip == 0:
                    rows.append([float(v) for v in line.split()])
                count += 1
        return np.array(rows)

    def read_traj(self, skip=1):
        """``format='txyz'``: (times [T], coordinates [T, 3]); ``format='xyz'``: coordinates [T, 3] only
        (``ves/utils.py:37-40,53-69``)."""
        if self.format not in ('txyz', 'xyz'):
            raise ValueError("Invalid format")
        data = self._rows(skip)
        if self.format == 'xyz':
            return data[:, 0:3]
        return data[:, 0], data[:, 1:4]

    def read_energies(self, skip=1):
        """(times [T], PE [T], KE [T]) from an energies.dat file"""
        data = self._rows(skip)
        return data[:, 0], data[:, 1], data[:, 2]
