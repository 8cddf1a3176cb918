"""ParaView export of a `.ttdb` series.

The reference's `data::export_hdf5(path, series)` (`tit/data/hdf5.cpp:312-338`)
writes `particles.xdmf` — an XDMF 3 temporal collection, one `Uniform` grid
`frame-NN` per frame with a `Polyvertex` topology, the geometry taken from the
array `r` and one node-centred `Attribute` per scalar / vector array (matrices
are skipped, `hdf5.cpp:178-181`) — next to `particles.h5` holding the heavy
data. There is no HDF5 library in this image, so the same XDMF document is
written here with its `DataItem`s pointing into one raw little-endian file
`particles.bin` (`Format="Binary"`, `Seek` = byte offset), which ParaView's
XDMF 3 reader opens just as well. Everything else — element order, names,
`Dimensions`, `NumberType` / `Precision` — follows the reference's writer.
"""
from __future__ import annotations

import math
import os
import xml.etree.ElementTree as ET

import numpy as np

from . import ttdb


def _number_type(dtype: np.dtype):
    """`NumberType`, `Precision` as `hdf5.cpp:258-297` assigns them (all integers are "Int")."""
    return ("Float" if dtype.kind == "f" else "Int"), str(dtype.itemsize)


def export_xdmf(path: str, series: "ttdb.Series") -> str:
    """Write `particles.xdmf` + `particles.bin` for all frames of `series` into the
    existing directory `path`; returns the path of the `.xdmf` file."""
    if not os.path.exists(path):
        raise FileNotFoundError("Directory does not exist!")
    if not os.path.isdir(path):
        raise NotADirectoryError("Path is not a directory!")
    xdmf_path, bin_name = os.path.join(path, "particles.xdmf"), "particles.bin"

    root = ET.Element("Xdmf", Version="3.0")
    domain = ET.SubElement(root, "Domain")
    collection = ET.SubElement(domain, "Grid", Name="TimeSeries", GridType="Collection", CollectionType="Temporal")

    frames = series.frames()
    padding = int(math.ceil(math.log10(max(1, len(frames)))))
    with open(os.path.join(path, bin_name), "wb") as heavy:

        def data_item(parent, values: np.ndarray):
            number_type, precision = _number_type(values.dtype)
            item = ET.SubElement(parent, "DataItem", Format="Binary", Dimensions=" ".join(str(d) for d in values.shape),
                                 NumberType=number_type, Precision=precision, Endian="Little", Seek=str(heavy.tell()))
            item.text = bin_name
            heavy.write(np.ascontiguousarray(values, dtype=values.dtype.newbyteorder("<")).tobytes())

        for index, frame in enumerate(frames):
            grid = ET.SubElement(collection, "Grid", Name=f"frame-{index:0{padding}d}" if padding else f"frame-{index}", GridType="Uniform")
            ET.SubElement(grid, "Time", Value=repr(float(frame.time)))
            positions = frame.find_array("r")
            if positions is None:
                raise KeyError("Positions array 'r' not found!")
            r = positions.read()
            ET.SubElement(grid, "Topology", TopologyType="Polyvertex", NumberOfElements=str(r.shape[0]))
            geometry = ET.SubElement(grid, "Geometry", GeometryType={1: "X", 2: "XY", 3: "XYZ"}[r.shape[1] if r.ndim > 1 else 1])
            data_item(geometry, r)
            for array in frame.arrays():
                _, rank, _ = ttdb.decode_type(array.type)
                if rank == ttdb.RANK_MATRIX:  # not exported by the reference either
                    continue
                attribute = ET.SubElement(grid, "Attribute", Name=array.name, Center="Node", AttributeType="Scalar" if rank == ttdb.RANK_SCALAR else "Vector")
                data_item(attribute, array.read())

    ET.indent(root)
    ET.ElementTree(root).write(xdmf_path, encoding="UTF-8", xml_declaration=True)
    return xdmf_path
