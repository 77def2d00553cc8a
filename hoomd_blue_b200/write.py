"""hoomd.write for this path (hoomd/write/gsd.py, hoomd/write/table.py): the GSD trajectory writer
(periodic, dynamic property categories, logged quantities as log/ chunks), its one-shot `GSD.write`,
and the Table writer for logged scalars."""
import sys

import numpy as np

from . import gsd_io
from .trigger import as_trigger


class Writer:
    pass


class GSD(Writer):
    """hoomd.write.GSD(trigger, filename, filter=All(), mode='ab', truncate=False, dynamic=None,
    logger=None) (hoomd/write/gsd.py:27-213, hoomd/GSDDumpWriter.cc). Frame 0 of the file holds every
    chunk; later frames hold configuration/*, the `dynamic` categories ('property' by default) and the
    logged quantities. filter must be All()."""

    def __init__(self, trigger, filename, filter=None, mode="ab", truncate=False, dynamic=None,
                 logger=None):
        from .filter import All
        if filter is not None and not isinstance(filter, All):
            raise NotImplementedError("this path writes filter.All() only")
        if mode not in ("wb", "xb", "ab"):
            raise ValueError("mode must be 'wb', 'xb' or 'ab'")
        self.trigger = as_trigger(trigger)
        self.filename = str(filename)
        self.mode = mode
        self.truncate = bool(truncate)
        self.dynamic = list(dynamic) if dynamic is not None else ["property"]
        for d in self.dynamic:
            if d.split("/")[0] not in ("property", "momentum", "attribute", "topology", "particles"):
                raise ValueError(f"unknown dynamic category {d!r}")
        self.logger = logger
        self._writer = None

    def _log_chunks(self):
        if self.logger is None:
            return None
        return {"/".join(ns): v for ns, v in self.logger._flat().items()}

    def _write(self, sim):
        state = sim.state
        snap = state.get_snapshot()  # (collective with a domain decomposition)
        log = self._log_chunks()
        if sim.device.communicator.rank != 0:
            return
        if self.truncate:
            gsd_io.GSDWriter(self.filename, "wb").append(
                gsd_io.frame_chunks(snap, sim.timestep, first=True, log=log))
            return
        if self._writer is None:
            self._writer = gsd_io.GSDWriter(self.filename, self.mode)
        dyn = set(d.split("/")[0] for d in self.dynamic)
        if "particles" in dyn:
            dyn |= {"property", "momentum", "attribute"}
        self._writer.append(gsd_io.frame_chunks(snap, sim.timestep, dynamic=dyn,
                                                first=self._writer.nframes == 0, log=log))

    def flush(self):
        pass  # every frame is complete on disk when _write returns

    @staticmethod
    def write(state, filename, filter=None, mode="wb", logger=None):
        """hoomd.write.GSD.write: one frame of the current state (hoomd/write/gsd.py:215-253)."""
        if mode not in ("wb", "xb"):
            raise ValueError("mode must be 'wb' or 'xb'")
        w = GSD(trigger=1, filename=filename, filter=filter, mode=mode, logger=logger)
        w._write(state._simulation)


class Table(Writer):
    """hoomd.write.Table(trigger, logger, output=stdout, header_sep='.', delimiter=' ', ...)
    (hoomd/write/table.py): a header line, then one line of the logged scalars per triggered step.
    Column names are the namespaces joined by header_sep, shortened to the last `max_header_len`
    components like the reference does when it is set."""

    def __init__(self, trigger, logger, output=None, header_sep=".", delimiter=" ", pretty=True,
                 max_precision=10, max_header_len=None):
        self.trigger = as_trigger(trigger)
        self.logger = logger
        self.output = output if output is not None else sys.stdout
        self.header_sep, self.delimiter = header_sep, delimiter
        self.max_precision = int(max_precision)
        self.max_header_len = max_header_len
        self._header_written = False

    def _write(self, sim):
        vals = self.logger._flat()
        if sim.device.communicator.rank != 0:
            return
        if not self._header_written:
            cols = []
            for ns in vals:
                parts = ns if self.max_header_len is None else ns[-self.max_header_len:]
                cols.append(self.header_sep.join(parts))
            self.output.write(self.delimiter.join(cols) + "\n")
            self._header_written = True
        row = []
        for v in vals.values():
            a = np.asarray(v)
            if a.ndim != 0:
                raise ValueError("hoomd.write.Table logs scalar quantities only")
            row.append(str(int(a)) if np.issubdtype(a.dtype, np.integer)
                       else f"{float(a):.{self.max_precision}g}")
        self.output.write(self.delimiter.join(row) + "\n")

    def flush(self):
        if hasattr(self.output, "flush"):
            self.output.flush()
