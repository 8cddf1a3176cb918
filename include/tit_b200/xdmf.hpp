// tit_b200/xdmf.hpp — ParaView export of a `.ttdb` series.
//
// The reference's data::export_hdf5(path, series) (tit/data/hdf5.cpp:312-338)
// writes `particles.xdmf` — an XDMF 3 temporal collection with one Uniform grid
// `frame-NN` per frame: Polyvertex topology, geometry from the array `r`, one
// node-centred Attribute per scalar / vector array (matrices skipped,
// hdf5.cpp:178-181) — beside `particles.h5` with the heavy data. Neither HDF5
// nor an XML library is available to this toolchain, so export_xdmf() prints the
// same document itself and points its DataItems into one raw little-endian file
// `particles.bin` (Format="Binary", Seek = byte offset), which ParaView's XDMF 3
// reader opens as well. titsolver_b200/xdmf.py is the same exporter in Python;
// tests/test_ttdb.py checks both against the frames they were made from.
#pragma once

#include <cmath>
#include <cstddef>
#include <cstdio>
#include <filesystem>
#include <fstream>
#include <string>

#include "data.hpp"

namespace tit::data {

namespace impl {
/// `Dimensions`, `NumberType` and `Precision` of an array, as hdf5.cpp:229-297 assigns them.
inline auto xdmf_item_attrs(Type type, std::size_t size) -> std::string {
  std::string dims = std::to_string(size);
  for (unsigned k = 0; k < static_cast<unsigned>(type.rank()); ++k) dims += " " + std::to_string(type.dim());
  const bool real = type.kind().id() == Kind::ID::float32 || type.kind().id() == Kind::ID::float64;
  return "Dimensions=\"" + dims + "\" NumberType=\"" + (real ? "Float" : "Int") + "\" Precision=\"" + std::to_string(type.kind().width()) + "\"";
}
inline auto xml_escape(const std::string& text) -> std::string {
  std::string out;
  for (const char ch : text) {
    switch (ch) {
      case '&': out += "&amp;"; break;
      case '<': out += "&lt;"; break;
      case '>': out += "&gt;"; break;
      case '"': out += "&quot;"; break;
      default: out += ch;
    }
  }
  return out;
}
}  // namespace impl

/// Write `particles.xdmf` and `particles.bin` for all frames of `series` into the
/// existing directory `path`; returns the path of the `.xdmf` file.
inline auto export_xdmf(const std::filesystem::path& path, SeriesView<const Storage> series) -> std::filesystem::path {
  if (!std::filesystem::exists(path)) throw Exception("Directory does not exist!");
  if (!std::filesystem::is_directory(path)) throw Exception("Path is not a directory!");
  const auto xdmf_path = path / "particles.xdmf";
  const std::string heavy_name = "particles.bin";
  std::ofstream heavy{path / heavy_name, std::ios::binary};
  std::ofstream xml{xdmf_path};
  if (!heavy || !xml) throw Exception("Unable to create the export files!");

  const auto frames = series.frames();
  const int padding = int(std::ceil(std::log10(double(std::max<std::size_t>(1, frames.size())))));
  std::size_t offset = 0;
  const auto data_item = [&](const ArrayView<const Storage>& array, const char* indent) {
    const auto bytes = array.read();
    xml << indent << "<DataItem Format=\"Binary\" " << impl::xdmf_item_attrs(array.type(), array.size()) << " Endian=\"Little\" Seek=\"" << offset << "\">" << heavy_name
        << "</DataItem>\n";
    heavy.write(reinterpret_cast<const char*>(bytes.data()), std::streamsize(bytes.size()));
    offset += bytes.size();
  };

  xml << "<?xml version=\"1.0\" encoding=\"UTF-8\"?>\n<Xdmf Version=\"3.0\">\n  <Domain>\n"
      << "    <Grid Name=\"TimeSeries\" GridType=\"Collection\" CollectionType=\"Temporal\">\n";
  std::size_t index = 0;
  for (const auto& frame : frames) {
    char name[48];
    std::snprintf(name, sizeof(name), "frame-%0*zu", padding, index++);
    char time[40];
    std::snprintf(time, sizeof(time), "%.17g", frame.time());
    const auto positions = frame.find_array("r");
    if (!positions) throw Exception("Positions array 'r' not found!");
    const std::size_t dim = positions->type().dim();
    xml << "      <Grid Name=\"" << name << "\" GridType=\"Uniform\">\n        <Time Value=\"" << time << "\" />\n"
        << "        <Topology TopologyType=\"Polyvertex\" NumberOfElements=\"" << positions->size() << "\" />\n"
        << "        <Geometry GeometryType=\"" << (dim == 1 ? "X" : dim == 2 ? "XY" : "XYZ") << "\">\n";
    data_item(*positions, "          ");
    xml << "        </Geometry>\n";
    for (const auto& array : frame.arrays()) {
      const auto type = array.type();
      if (type.rank() == Rank::matrix) continue;  // not exported by the reference either
      xml << "        <Attribute Name=\"" << impl::xml_escape(array.name()) << "\" Center=\"Node\" AttributeType=\"" << (type.rank() == Rank::scalar ? "Scalar" : "Vector") << "\">\n";
      data_item(array, "          ");
      xml << "        </Attribute>\n";
    }
    xml << "      </Grid>\n";
  }
  xml << "    </Grid>\n  </Domain>\n</Xdmf>\n";
  if (!heavy || !xml) throw Exception("Unable to write the export files!");
  return xdmf_path;
}

}  // namespace tit::data
