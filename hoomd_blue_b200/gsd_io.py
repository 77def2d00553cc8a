"""Minimal GSD (hoomd schema) frame reader/writer — placeholder until the file-format row lands."""


def read_frame(filename, frame=-1):
    raise NotImplementedError("GSD I/O is not implemented yet")
