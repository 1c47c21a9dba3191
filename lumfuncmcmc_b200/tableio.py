"""Minimal column table with the two on-disk formats the drivers use (astropy is not in this image).

Read : whitespace-separated ASCII with one header line of column names (what ``Table.read(fn, format='ascii')``
       accepts for the reference's catalogues: columns ``Field z ID <line>_flux <line>_flux_e``,
       reference run_lumfuncmcmc.py:165-179); also reads back the two-line fixed-width files written below.
Write: ``ascii.fixed_width_two_line`` -- a header line, a line of dashes, then rows, all columns padded to a
       common width (reference run_lumfuncmcmc.py:297-330).
Only what the drivers and ``add_fitinfo_to_table`` touch is implemented.
"""
import numpy as np


class Row:
    def __init__(self, table, index):
        self._t, self._i = table, index

    def __len__(self):
        return len(self._t.colnames)

    def _name(self, key):
        return self._t.colnames[key] if isinstance(key, (int, np.integer)) else key

    def __getitem__(self, key):
        return self._t._cols[self._name(key)][self._i]

    def __setitem__(self, key, value):
        self._t._cols[self._name(key)][self._i] = value


class Table:
    def __init__(self, data=None, names=None, dtype=None):
        self.colnames = list(names) if names is not None else []
        self._cols = {}
        if data is None:
            dtype = dtype or ['f8'] * len(self.colnames)
            for n, dt in zip(self.colnames, dtype):
                self._cols[n] = [] if str(dt).startswith(('S', 'U', 'str')) else np.zeros(0, dtype=np.float64)
        else:
            if isinstance(data, np.ndarray) and data.ndim == 2:
                cols = [data[:, j] for j in range(data.shape[1])]
            else:
                cols = [np.asarray(c) for c in data]
            if not self.colnames:
                self.colnames = ['col%d' % j for j in range(len(cols))]
            for n, c in zip(self.colnames, cols):
                self._cols[n] = np.array(c)

    # -- access -------------------------------------------------------------------------------
    @property
    def columns(self):
        return self.colnames

    def __len__(self):
        return len(self._cols[self.colnames[0]]) if self.colnames else 0

    def __getitem__(self, key):
        if isinstance(key, str):
            return np.asarray(self._cols[key])
        if isinstance(key, (int, np.integer)):
            n = len(self)
            return Row(self, key if key >= 0 else n + key)
        raise KeyError(key)

    def add_row(self, values):
        for n, v in zip(self.colnames, values):
            col = self._cols[n]
            if isinstance(col, list):
                col.append(v)
            else:
                self._cols[n] = np.append(col, v)

    def as_array(self):
        return np.column_stack([np.asarray(self._cols[n], dtype=np.float64) for n in self.colnames])

    def __repr__(self):
        return self._render({})

    # -- formats ------------------------------------------------------------------------------
    def _render(self, formats):
        def fmt(name, v):
            f = formats.get(name)
            if f is not None:
                return (f % v) if '%' in f else format(v, f)
            if isinstance(v, (bytes, np.bytes_)):
                return v.decode()
            if isinstance(v, (float, np.floating)):
                return repr(float(v))
            return str(v)
        cells = [[fmt(n, v) for v in self._cols[n]] for n in self.colnames]
        widths = [max([len(n)] + [len(c) for c in col]) for n, col in zip(self.colnames, cells)]
        lines = [' '.join(n.rjust(w) for n, w in zip(self.colnames, widths)),
                 ' '.join('-' * w for w in widths)]
        for i in range(len(self)):
            lines.append(' '.join(col[i].rjust(w) for col, w in zip(cells, widths)))
        return '\n'.join(lines)

    def write(self, path, format='ascii.fixed_width_two_line', formats=None, overwrite=True):
        if format not in ('ascii.fixed_width_two_line', 'ascii'):
            raise ValueError("unsupported table format %r" % format)
        import os
        if os.path.exists(path) and not overwrite:
            raise OSError("%s exists" % path)
        with open(path, 'w') as fh:
            fh.write(self._render(formats or {}) + '\n')

    @classmethod
    def read(cls, path, format='ascii'):
        with open(path) as fh:
            lines = [ln.rstrip('\n') for ln in fh if ln.strip() and not ln.lstrip().startswith('#')]
        if len(lines) > 1 and set(lines[1].strip()) <= set('- '):
            # two-line fixed width: column spans come from the dash line (names may contain spaces, e.g. 'Ln Prob')
            spans, pos = [], 0
            for tok in lines[1].split(' '):
                if tok:
                    spans.append((pos, pos + len(tok)))
                pos += len(tok) + 1
            names = [lines[0][a:b].strip() for a, b in spans]
            rows = [[ln[a:b].strip() for a, b in spans] for ln in lines[2:]]
        else:
            names = lines[0].split()
            rows = [ln.split() for ln in lines[1:]]
        cols = []
        for j in range(len(names)):
            raw = [r[j] for r in rows]
            try:
                cols.append(np.array([int(v) for v in raw]))
            except ValueError:
                try:
                    cols.append(np.array([float(v) for v in raw]))
                except ValueError:
                    cols.append(np.array(raw))
        return cls(cols, names=names)
