// tit_b200/data.hpp — the `.ttdb` particle storage behind ParticleArray::write
// (the data format on the output side of the particle step).
//
// Keeps the reference's public surface — tit::data::Storage and its
// SeriesView / FrameView / ArrayView handles, data::Type / data::Kind and
// data::type_of (tit/data/storage.hpp:46-520, tit/data/type.hpp:26-236) — and its
// on-disk format, so a database written here opens in the reference's tools and
// GUI and vice versa:
//   * an SQLite file with the tables Settings / DataSeries / DataFrames /
//     DataArrays (tit/data/storage.cpp:33-64), oldest series evicted beyond
//     `max_series` (storage.cpp:130-143), cascaded deletes;
//   * DataArrays.type = (kind + 1) | rank << 8 | dim << 16 (type.hpp:176-189),
//     DataArrays.size = number of elements, DataArrays.data = the packed
//     little-endian elements (core/serialization.hpp:57-128: a Vec is its Dim
//     numbers, a Mat its Dim rows) as Zstandard frames (storage.cpp:351-372).
//
// Built differently from the reference: no build-time dependency on sqlite3.h /
// zstd.h (neither is installed next to this toolchain) — the two runtime
// libraries are bound with dlopen on first use, through a three-function query
// helper instead of a statement class per call site, and arrays are
// (de)compressed in one piece since the facade's columns are contiguous and
// packed. C++20, header-only; link with -ldl on old glibc.
#pragma once

#include <dlfcn.h>

#include <array>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <filesystem>
#include <initializer_list>
#include <optional>
#include <ranges>
#include <span>
#include <string>
#include <string_view>
#include <type_traits>
#include <utility>
#include <vector>

#include "core.hpp"

namespace tit {

using float32_t = float;

namespace data {

// ~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~
// Type descriptors (tit/data/type.hpp).

/// Scalar kind of an array element (type.hpp:26-110).
class Kind final {
public:
  enum class ID : std::uint8_t { int8, uint8, int16, uint16, int32, uint32, int64, uint64, float32, float64, unknown_ };
  constexpr explicit Kind(ID id) : id_{id} {
    if (id_ >= ID::unknown_) throw Exception("Invalid data kind ID: " + std::to_string(static_cast<unsigned>(id_)) + ".");
  }
  constexpr auto id() const noexcept -> ID { return id_; }
  /// Width in bytes: the ids come in (signed, unsigned) pairs of doubling width, then the two floats.
  constexpr auto width() const noexcept -> std::size_t {
    const auto k = static_cast<unsigned>(id_);
    return k < 8 ? std::size_t{1} << (k / 2) : (id_ == ID::float32 ? 4 : 8);
  }
  constexpr auto name() const noexcept -> const char* {
    constexpr const char* names[] = {"int8_t", "uint8_t", "int16_t", "uint16_t", "int32_t", "uint32_t", "int64_t", "uint64_t", "float32_t", "float64_t"};
    return names[static_cast<unsigned>(id_)];
  }
  constexpr auto operator==(const Kind&) const noexcept -> bool = default;
private:
  ID id_;
};

namespace impl {
template<class Val> inline constexpr auto kind_id_of = Kind::ID::unknown_;
template<> inline constexpr auto kind_id_of<std::int8_t> = Kind::ID::int8;
template<> inline constexpr auto kind_id_of<std::uint8_t> = Kind::ID::uint8;
template<> inline constexpr auto kind_id_of<std::int16_t> = Kind::ID::int16;
template<> inline constexpr auto kind_id_of<std::uint16_t> = Kind::ID::uint16;
template<> inline constexpr auto kind_id_of<std::int32_t> = Kind::ID::int32;
template<> inline constexpr auto kind_id_of<std::uint32_t> = Kind::ID::uint32;
template<> inline constexpr auto kind_id_of<std::int64_t> = Kind::ID::int64;
template<> inline constexpr auto kind_id_of<std::uint64_t> = Kind::ID::uint64;
template<> inline constexpr auto kind_id_of<float32_t> = Kind::ID::float32;
template<> inline constexpr auto kind_id_of<float64_t> = Kind::ID::float64;
}  // namespace impl

template<class Val>
concept known_kind_of = std::is_object_v<Val> && (impl::kind_id_of<std::remove_cv_t<Val>> < Kind::ID::unknown_);
template<known_kind_of Val> inline constexpr Kind kind_of{impl::kind_id_of<std::remove_cv_t<Val>>};

enum class Rank : std::uint8_t { scalar, vector, matrix, count_ };

/// Element type of an array: kind, rank, dimension (type.hpp:143-216).
class Type final {
public:
  constexpr explicit Type(Kind kind) : Type{kind, Rank::scalar, 1} {}
  constexpr explicit Type(Kind kind, Rank rank, std::uint8_t dim) : kind_{kind}, rank_{rank}, dim_{dim} {
    if (rank >= Rank::count_) throw Exception("Invalid data type rank: " + std::to_string(static_cast<unsigned>(rank)) + ".");
    if (dim == 0) throw Exception("Dimensionality must be positive, but is 0.");
    if (rank == Rank::scalar && dim != 1) throw Exception("Dimensionality of a scalar must be 1, but is " + std::to_string(unsigned{dim}) + ".");
  }
  /// From the integer stored in DataArrays.type.
  constexpr explicit Type(std::uint32_t id)
      : Type{Kind{static_cast<Kind::ID>((id - 1) & 0xFF)}, static_cast<Rank>((id >> 8) & 0xFF), static_cast<std::uint8_t>((id >> 16) & 0xFF)} {}
  constexpr auto id() const noexcept -> std::uint32_t {
    return (static_cast<std::uint32_t>(kind_.id()) + 1) | static_cast<std::uint32_t>(rank_) << 8 | static_cast<std::uint32_t>(dim_) << 16;
  }
  constexpr auto kind() const noexcept -> Kind { return kind_; }
  constexpr auto rank() const noexcept -> Rank { return rank_; }
  constexpr auto dim() const noexcept -> std::size_t { return dim_; }
  /// Bytes per element: kind width × dim^rank.
  constexpr auto width() const noexcept -> std::size_t {
    std::size_t w = kind_.width();
    for (unsigned r = 0; r < static_cast<unsigned>(rank_); ++r) w *= dim_;
    return w;
  }
  auto name() const -> std::string {
    if (rank_ == Rank::scalar) return kind_.name();
    return std::string{rank_ == Rank::vector ? "Vec<" : "Mat<"} + kind_.name() + ", " + std::to_string(dim()) + ">";
  }
  constexpr auto operator==(const Type&) const noexcept -> bool = default;
private:
  Kind kind_;
  Rank rank_;
  std::uint8_t dim_;
};

namespace impl {
template<class Val> struct type_of_t;
template<known_kind_of Val> struct type_of_t<Val> { static constexpr Type value{kind_of<Val>}; };
template<known_kind_of Num, std::size_t Dim> struct type_of_t<Vec<Num, Dim>> { static constexpr Type value{kind_of<Num>, Rank::vector, Dim}; };
template<known_kind_of Num, std::size_t Dim> struct type_of_t<Mat<Num, Dim>> { static constexpr Type value{kind_of<Num>, Rank::matrix, Dim}; };
}  // namespace impl
template<class Val>
concept known_type_of = requires { impl::type_of_t<std::remove_cv_t<Val>>::value.id(); };
template<known_type_of Val> inline constexpr Type type_of = impl::type_of_t<std::remove_cv_t<Val>>::value;

// ~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~
// Runtime bindings of libsqlite3 / libzstd.

namespace impl {

inline auto open_lib(std::initializer_list<const char*> names) -> void* {
  for (const char* n : names)
    if (void* h = dlopen(n, RTLD_NOW | RTLD_GLOBAL)) return h;
  throw Exception(std::string{"cannot load "} + *names.begin() + ": " + dlerror());
}
template<class Fn>
void bind_sym(void* lib, const char* name, Fn*& fn) {
  fn = reinterpret_cast<Fn*>(dlsym(lib, name));
  if (fn == nullptr) throw Exception(std::string{"missing symbol "} + name);
}

/// The dozen entry points of the SQLite C API this header uses.
struct Sqlite final {
  using db_t = void;
  using stmt_t = void;
  int (*open_v2)(const char*, db_t**, int, const char*);
  int (*close_v2)(db_t*);
  int (*exec)(db_t*, const char*, int (*)(void*, int, char**, char**), void*, char**);
  int (*prepare_v2)(db_t*, const char*, int, stmt_t**, const char**);
  int (*bind_int64)(stmt_t*, int, long long);
  int (*bind_double)(stmt_t*, int, double);
  int (*bind_text)(stmt_t*, int, const char*, int, void (*)(void*));
  int (*bind_blob64)(stmt_t*, int, const void*, unsigned long long, void (*)(void*));
  int (*step)(stmt_t*);
  int (*finalize)(stmt_t*);
  int (*column_type)(stmt_t*, int);
  long long (*column_int64)(stmt_t*, int);
  double (*column_double)(stmt_t*, int);
  const unsigned char* (*column_text)(stmt_t*, int);
  const void* (*column_blob)(stmt_t*, int);
  int (*column_bytes)(stmt_t*, int);
  const char* (*errmsg)(db_t*);
  long long (*last_insert_rowid)(db_t*);
  const char* (*db_filename)(db_t*, const char*);
  static constexpr int open_readonly = 1, open_readwrite = 2, open_create = 4, row = 100, done = 101, null_type = 5;
  static auto api() -> const Sqlite& {
    static const Sqlite s = [] {
      Sqlite q{};
      void* lib = open_lib({"libsqlite3.so.0", "libsqlite3.so"});
#define TIT_B200_SQL(f) bind_sym(lib, "sqlite3_" #f, q.f)
      TIT_B200_SQL(open_v2); TIT_B200_SQL(close_v2); TIT_B200_SQL(exec); TIT_B200_SQL(prepare_v2); TIT_B200_SQL(bind_int64);
      TIT_B200_SQL(bind_double); TIT_B200_SQL(bind_text); TIT_B200_SQL(bind_blob64); TIT_B200_SQL(step); TIT_B200_SQL(finalize);
      TIT_B200_SQL(column_type); TIT_B200_SQL(column_int64); TIT_B200_SQL(column_double); TIT_B200_SQL(column_text);
      TIT_B200_SQL(column_blob); TIT_B200_SQL(column_bytes); TIT_B200_SQL(errmsg); TIT_B200_SQL(last_insert_rowid); TIT_B200_SQL(db_filename);
#undef TIT_B200_SQL
      return q;
    }();
    return s;
  }
};

/// Zstandard: one-shot compression, streaming decompression (the reference writes
/// streamed frames without a content size, core/zstd.cpp:53-100).
struct Zstd final {
  struct InBuf { const void* src; std::size_t size, pos; };
  struct OutBuf { void* dst; std::size_t size, pos; };
  std::size_t (*compressBound)(std::size_t);
  std::size_t (*compress)(void*, std::size_t, const void*, std::size_t, int);
  unsigned (*isError)(std::size_t);
  const char* (*getErrorName)(std::size_t);
  void* (*createDStream)();
  std::size_t (*freeDStream)(void*);
  std::size_t (*decompressStream)(void*, OutBuf*, InBuf*);
  static auto api() -> const Zstd& {
    static const Zstd s = [] {
      Zstd z{};
      void* lib = open_lib({"libzstd.so.1", "libzstd.so"});
#define TIT_B200_ZSTD(f) bind_sym(lib, "ZSTD_" #f, z.f)
      TIT_B200_ZSTD(compressBound); TIT_B200_ZSTD(compress); TIT_B200_ZSTD(isError); TIT_B200_ZSTD(getErrorName);
      TIT_B200_ZSTD(createDStream); TIT_B200_ZSTD(freeDStream); TIT_B200_ZSTD(decompressStream);
#undef TIT_B200_ZSTD
      return z;
    }();
    return s;
  }
};

inline auto zstd_pack(std::span<const std::byte> raw) -> std::vector<std::byte> {
  const auto& z = Zstd::api();
  std::vector<std::byte> out(z.compressBound(raw.size()));
  const std::size_t n = z.compress(out.data(), out.size(), raw.data(), raw.size(), /*level=*/3);
  if (z.isError(n) != 0) throw Exception(std::string{"ZSTD compression failed: "} + z.getErrorName(n) + ".");
  out.resize(n);
  return out;
}

/// Decompress the concatenated frames of `packed` into exactly `raw.size()` bytes.
inline void zstd_unpack(std::span<const std::byte> packed, std::span<std::byte> raw) {
  const auto& z = Zstd::api();
  void* ds = z.createDStream();
  if (ds == nullptr) throw Exception("ZSTD decompression failed: no context.");
  Zstd::InBuf in{packed.data(), packed.size(), 0};
  Zstd::OutBuf out{raw.data(), raw.size(), 0};
  std::size_t status = 0;
  while (in.pos < in.size) {
    const std::size_t out_before = out.pos, in_before = in.pos;
    status = z.decompressStream(ds, &out, &in);
    if (z.isError(status) != 0) {
      const std::string why = z.getErrorName(status);
      z.freeDStream(ds);
      throw Exception("ZSTD decompression failed: " + why + ".");
    }
    if (out.pos == out_before && in.pos == in_before) break;  // output full with input left over
  }
  z.freeDStream(ds);
  if (status != 0 || in.pos != in.size) throw Exception(in.pos != in.size ? "ZSTD decompression failed: data size mismatch." : "ZSTD decompression failed: truncated frame.");
  if (out.pos != out.size) throw Exception("ZSTD decompression failed: data size mismatch.");
}

/// One database connection with a variadic query helper.
class Db final {
public:
  Db(const std::filesystem::path& path, bool read_only) {
    const auto& q = Sqlite::api();
    const int flags = read_only ? Sqlite::open_readonly : Sqlite::open_readwrite | Sqlite::open_create;
    if (q.open_v2(path.c_str(), &db_, flags, nullptr) != 0) {
      const std::string why = db_ != nullptr ? q.errmsg(db_) : "out of memory";
      q.close_v2(db_);
      throw Exception("SQLite: " + why + " ('" + path.string() + "').");
    }
    // A file that is not a database is only noticed by the first statement.
    if (q.exec(db_, "SELECT COUNT(*) FROM sqlite_master", nullptr, nullptr, nullptr) != 0) {
      const std::string why = q.errmsg(db_);
      q.close_v2(db_);
      throw Exception("SQLite: " + why + " ('" + path.string() + "').");
    }
  }
  Db(const Db&) = delete;
  auto operator=(const Db&) -> Db& = delete;
  ~Db() { Sqlite::api().close_v2(db_); }

  auto path() const -> std::filesystem::path {
    const char* p = Sqlite::api().db_filename(db_, "main");
    return p != nullptr ? std::filesystem::path{p} : std::filesystem::path{};
  }
  void script(const char* sql) const {
    if (Sqlite::api().exec(db_, sql, nullptr, nullptr, nullptr) != 0) fail_();
  }
  auto last_insert_row_id() const -> std::int64_t { return Sqlite::api().last_insert_rowid(db_); }

  /// Run `sql` with `args` bound to ?1, ?2, ...; `row(stmt)` is called per result row
  /// and returns whether to continue.
  template<class Row, class... Args>
  void each(const char* sql, Row&& row, const Args&... args) const {
    const auto& q = Sqlite::api();
    Sqlite::stmt_t* st = nullptr;
    if (q.prepare_v2(db_, sql, -1, &st, nullptr) != 0) fail_();
    int idx = 0;
    const bool bound = ((bind_(st, ++idx, args) == 0) && ...);
    int rc = bound ? Sqlite::row : 1;
    while (bound && (rc = q.step(st)) == Sqlite::row)
      if (!row(st)) { rc = Sqlite::done; break; }
    const std::string why = rc == Sqlite::done ? std::string{} : std::string{q.errmsg(db_)};
    q.finalize(st);
    if (rc != Sqlite::done) throw Exception("SQLite: " + why + ".");
  }
  template<class... Args>
  void run(const char* sql, const Args&... args) const {
    each(sql, [](Sqlite::stmt_t*) { return true; }, args...);
  }
  /// First column of the first row, if any.
  template<class T, class... Args>
  auto first(const char* sql, const Args&... args) const -> std::optional<T> {
    std::optional<T> result;
    each(sql, [&result](Sqlite::stmt_t* st) { result = column_<T>(st); return false; }, args...);
    return result;
  }
  /// First column of every row.
  template<class T, class... Args>
  auto all(const char* sql, const Args&... args) const -> std::vector<T> {
    std::vector<T> result;
    each(sql, [&result](Sqlite::stmt_t* st) { result.push_back(column_<T>(st)); return true; }, args...);
    return result;
  }

private:
  [[noreturn]] void fail_() const { throw Exception(std::string{"SQLite: "} + Sqlite::api().errmsg(db_) + "."); }
  template<class T>
  static auto bind_(Sqlite::stmt_t* st, int idx, const T& v) -> int {
    const auto& q = Sqlite::api();
    if constexpr (std::is_enum_v<T>) return q.bind_int64(st, idx, static_cast<long long>(v));
    else if constexpr (std::is_integral_v<T>) return q.bind_int64(st, idx, static_cast<long long>(v));
    else if constexpr (std::is_floating_point_v<T>) return q.bind_double(st, idx, v);
    else if constexpr (std::is_convertible_v<T, std::string_view>) {
      const std::string_view s{v};
      return q.bind_text(st, idx, s.data(), static_cast<int>(s.size()), reinterpret_cast<void (*)(void*)>(-1) /*copy*/);
    } else {
      const std::span<const std::byte> b{v};
      return q.bind_blob64(st, idx, b.data() != nullptr ? static_cast<const void*>(b.data()) : "", b.size(), nullptr /*static*/);
    }
  }
  template<class T>
  static auto column_(Sqlite::stmt_t* st) -> T {
    const auto& q = Sqlite::api();
    if constexpr (std::is_enum_v<T> || std::is_integral_v<T>) return static_cast<T>(q.column_int64(st, 0));
    else if constexpr (std::is_floating_point_v<T>) return q.column_double(st, 0);
    else if constexpr (std::is_same_v<T, std::string>) {
      const auto* s = q.column_text(st, 0);
      return s != nullptr ? std::string{reinterpret_cast<const char*>(s), static_cast<std::size_t>(q.column_bytes(st, 0))} : std::string{};
    } else {
      const auto* b = static_cast<const std::byte*>(q.column_blob(st, 0));
      return b != nullptr ? T(b, b + q.column_bytes(st, 0)) : T{};
    }
  }
  Sqlite::db_t* db_ = nullptr;
};

}  // namespace impl

// ~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~
// Handles (tit/data/storage.hpp:32-310).

enum class SeriesID : std::int64_t {};
enum class FrameID : std::int64_t {};
enum class ArrayID : std::int64_t {};

class Storage;
template<class S>
concept storage = std::same_as<std::remove_const_t<S>, Storage>;

namespace impl {
/// What the three handles share: the storage pointer, the row id, comparison by id.
template<class S, class ID>
class Handle {
public:
  constexpr Handle() noexcept = default;
  constexpr Handle(S& s, ID id) noexcept : storage_{&s}, id_{id} {}
  constexpr auto storage() const noexcept -> S& { return *storage_; }
  constexpr auto id() const noexcept -> ID { return id_; }
  constexpr operator ID() const noexcept { return id_; }  // NOLINT: the reference's handles convert to their ids too
  friend constexpr auto operator==(const Handle& a, const Handle& b) noexcept -> bool { return a.id_ == b.id_; }
  friend constexpr auto operator==(const Handle& a, ID b) noexcept -> bool { return a.id_ == b; }
protected:
  S* storage_ = nullptr;
  ID id_{0};
};
}  // namespace impl

/// One array of a frame (storage.hpp:46-144).
template<storage S>
class ArrayView final : public impl::Handle<S, ArrayID> {
public:
  using impl::Handle<S, ArrayID>::Handle;
  template<storage O>
    requires (!std::same_as<O, S> && std::convertible_to<O&, S&>)
  constexpr ArrayView(ArrayView<O> o) noexcept : impl::Handle<S, ArrayID>{o.storage(), o.id()} {}  // NOLINT
  auto name() const -> std::string { return this->storage().array_name(this->id()); }
  auto type() const -> Type { return this->storage().array_type(this->id()); }
  auto size() const -> std::size_t { return this->storage().array_size(this->id()); }
  void write(Type type, std::span<const std::byte> bytes) const { this->storage().array_write(this->id(), type, bytes); }
  template<std::ranges::contiguous_range Range>
    requires std::ranges::sized_range<Range> && known_type_of<std::ranges::range_value_t<Range>>
  void write(Range&& values) const { this->storage().array_write(this->id(), std::forward<Range>(values)); }
  void read(std::span<std::byte> bytes) const { this->storage().array_read(this->id(), bytes); }
  auto read() const -> std::vector<std::byte> { return this->storage().array_read(this->id()); }
  template<known_type_of Val>
  auto read() const -> std::vector<Val> { return this->storage().template array_read<Val>(this->id()); }
};
template<class S> ArrayView(S&, ArrayID) -> ArrayView<S>;

/// One time frame of a series (storage.hpp:148-226).
template<storage S>
class FrameView final : public impl::Handle<S, FrameID> {
public:
  using impl::Handle<S, FrameID>::Handle;
  template<storage O>
    requires (!std::same_as<O, S> && std::convertible_to<O&, S&>)
  constexpr FrameView(FrameView<O> o) noexcept : impl::Handle<S, FrameID>{o.storage(), o.id()} {}  // NOLINT
  auto time() const -> float64_t { return this->storage().frame_time(this->id()); }
  auto num_arrays() const -> std::size_t { return this->storage().frame_num_arrays(this->id()); }
  auto arrays() const { return this->storage().frame_arrays(this->id()); }
  auto find_array(std::string_view name) const { return this->storage().frame_find_array(this->id(), name); }
  auto create_array(std::string_view name) const -> ArrayView<S>
    requires (!std::is_const_v<S>)
  { return this->storage().frame_create_array(this->id(), name); }
};
template<class S> FrameView(S&, FrameID) -> FrameView<S>;

/// One series = one run (storage.hpp:229-310).
template<storage S>
class SeriesView final : public impl::Handle<S, SeriesID> {
public:
  using impl::Handle<S, SeriesID>::Handle;
  template<storage O>
    requires (!std::same_as<O, S> && std::convertible_to<O&, S&>)
  constexpr SeriesView(SeriesView<O> o) noexcept : impl::Handle<S, SeriesID>{o.storage(), o.id()} {}  // NOLINT
  auto name() const -> std::string { return this->storage().series_name(this->id()); }
  auto num_frames() const -> std::size_t { return this->storage().series_num_frames(this->id()); }
  auto frame(std::size_t index) const { return this->storage().series_frame(this->id(), index); }
  auto frames() const { return this->storage().series_frames(this->id()); }
  auto last_frame() const { return this->storage().series_last_frame(this->id()); }
  auto create_frame(float64_t time) const -> FrameView<S>
    requires (!std::is_const_v<S>)
  { return this->storage().series_create_frame(this->id(), time); }
};
template<class S> SeriesView(S&, SeriesID) -> SeriesView<S>;

// ~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~~
/// The storage (tit/data/storage.hpp:314-520, storage.cpp).
class Storage final {
  template<class T>
  static auto need_(const std::optional<T>& v, const std::string& what) -> const T& {
    if (!v) throw Exception(what);
    return *v;
  }
  template<template<class> class View, class Self, class ID>
  static auto views_(Self& self, const std::vector<ID>& ids) -> std::vector<View<Self>> {
    std::vector<View<Self>> out;
    out.reserve(ids.size());
    for (const ID id : ids) out.emplace_back(self, id);
    return out;
  }
  template<class Self>
  static auto find_(Self& self, FrameID id, std::string_view name) -> std::optional<ArrayView<Self>> {
    if (const auto a = self.frame_find_array_id(id, name)) return ArrayView<Self>{self, *a};
    return std::nullopt;
  }

public:
  /// Open or create the database file; ":memory:" gives a transient one.
  explicit Storage(const std::filesystem::path& path, bool read_only = false) : db_{path, read_only} {
    if (read_only) return;
    db_.script(R"SQL(
      PRAGMA journal_mode = WAL;
      PRAGMA foreign_keys = ON;
      CREATE TABLE IF NOT EXISTS Settings (id INTEGER PRIMARY KEY CHECK (id = 0), max_series INTEGER) STRICT;
      INSERT OR IGNORE INTO Settings (id, max_series) VALUES (0, 5);
      CREATE TABLE IF NOT EXISTS DataSeries (id INTEGER PRIMARY KEY AUTOINCREMENT, name TEXT NOT NULL) STRICT;
      CREATE TABLE IF NOT EXISTS DataFrames (
        id INTEGER PRIMARY KEY AUTOINCREMENT, series_id INTEGER NOT NULL, time REAL NOT NULL,
        FOREIGN KEY (series_id) REFERENCES DataSeries(id) ON DELETE CASCADE) STRICT;
      CREATE TABLE IF NOT EXISTS DataArrays (
        id INTEGER PRIMARY KEY AUTOINCREMENT, frame_id INTEGER NOT NULL, name TEXT NOT NULL,
        type INTEGER, size INTEGER, data BLOB,
        FOREIGN KEY (frame_id) REFERENCES DataFrames(id) ON DELETE CASCADE) STRICT;
    )SQL");
  }

  auto path() const -> std::filesystem::path { return db_.path(); }

  // ~~ series ~~
  auto max_series() const -> std::size_t { return need_(db_.first<std::size_t>("SELECT max_series FROM Settings"), "Unable to get maximum number of series!"); }
  /// Lower the cap: the oldest series beyond it go (storage.cpp:77-92).
  void set_max_series(std::size_t value) {
    if (value == 0) throw Exception("Maximum number of series must be positive!");
    db_.run("UPDATE Settings SET max_series = ?", value);
    if (const auto n = num_series(); n > value) db_.run("DELETE FROM DataSeries WHERE id IN (SELECT id FROM DataSeries ORDER BY id ASC LIMIT ?)", n - value);
  }
  auto num_series() const -> std::size_t { return need_(db_.first<std::size_t>("SELECT COUNT(*) FROM DataSeries"), "Unable to count series!"); }
  auto series_id(std::size_t index) const -> SeriesID {
    return need_(db_.first<SeriesID>("SELECT id FROM DataSeries ORDER BY id ASC LIMIT 1 OFFSET ?", index), "Series index '" + std::to_string(index) + "' out of bounds.");
  }
  auto series_ids() const -> std::vector<SeriesID> { return db_.all<SeriesID>("SELECT id FROM DataSeries ORDER BY id ASC"); }
  auto last_series_id() const -> SeriesID { return need_(db_.first<SeriesID>("SELECT id FROM DataSeries ORDER BY id DESC LIMIT 1"), "Unable to get last series!"); }
  /// New series; at the cap the oldest one is evicted first (storage.cpp:130-143).
  auto create_series_id(std::string_view name = "") -> SeriesID {
    if (num_series() >= max_series()) db_.run("DELETE FROM DataSeries WHERE id IN (SELECT id FROM DataSeries ORDER BY id ASC LIMIT 1)");
    db_.run("INSERT INTO DataSeries (name) VALUES (?)", name);
    return SeriesID{db_.last_insert_row_id()};
  }
  void delete_series(SeriesID id) { db_.run("DELETE FROM DataSeries WHERE id = ?", id); }
  auto check_series(SeriesID id) const -> bool { return db_.first<SeriesID>("SELECT id FROM DataSeries WHERE id = ?", id).has_value(); }
  auto series_name(SeriesID id) const -> std::string { return need_(db_.first<std::string>("SELECT name FROM DataSeries WHERE id = ?", id), "Unable to get series name!"); }

  auto series(std::size_t index) { return SeriesView{*this, series_id(index)}; }
  auto series(std::size_t index) const { return SeriesView{*this, series_id(index)}; }
  auto series() { return views_<SeriesView>(*this, series_ids()); }
  auto series() const { return views_<SeriesView>(*this, series_ids()); }
  auto last_series() { return SeriesView{*this, last_series_id()}; }
  auto last_series() const { return SeriesView{*this, last_series_id()}; }
  auto create_series(std::string_view name = "") -> SeriesView<Storage> { return SeriesView{*this, create_series_id(name)}; }

  // ~~ frames ~~
  auto series_num_frames(SeriesID id) const -> std::size_t { return need_(db_.first<std::size_t>("SELECT COUNT(*) FROM DataFrames WHERE series_id = ?", id), "Unable to count frames!"); }
  auto series_frame_id(SeriesID id, std::size_t index) const -> FrameID {
    return need_(db_.first<FrameID>("SELECT id FROM DataFrames WHERE series_id = ? ORDER BY id ASC LIMIT 1 OFFSET ?", id, index), "Frame index '" + std::to_string(index) + "' out of bounds.");
  }
  auto series_frame_ids(SeriesID id) const -> std::vector<FrameID> { return db_.all<FrameID>("SELECT id FROM DataFrames WHERE series_id = ? ORDER BY id ASC", id); }
  auto series_last_frame_id(SeriesID id) const -> FrameID {
    return need_(db_.first<FrameID>("SELECT id FROM DataFrames WHERE series_id = ? ORDER BY id DESC LIMIT 1", id), "Unable to get last time step!");
  }
  /// Frames of a series carry increasing times (asserted by the reference, storage.cpp:224-226; checked here).
  auto series_create_frame_id(SeriesID id, float64_t time) -> FrameID {
    if (!check_series(id)) throw Exception("Invalid series ID!");
    if (const auto last = db_.first<float64_t>("SELECT time FROM DataFrames WHERE series_id = ? ORDER BY id DESC LIMIT 1", id); last && !(time > *last))
      throw Exception("Frame time must be greater than the last frame time!");
    db_.run("INSERT INTO DataFrames (series_id, time) VALUES (?, ?)", id, time);
    return FrameID{db_.last_insert_row_id()};
  }
  void delete_frame(FrameID id) { db_.run("DELETE FROM DataFrames WHERE id = ?", id); }
  auto check_frame(FrameID id) const -> bool { return db_.first<FrameID>("SELECT id FROM DataFrames WHERE id = ?", id).has_value(); }
  auto frame_time(FrameID id) const -> float64_t { return need_(db_.first<float64_t>("SELECT time FROM DataFrames WHERE id = ?", id), "Unable to get frame time!"); }

  auto series_frame(SeriesID id, std::size_t index) { return FrameView{*this, series_frame_id(id, index)}; }
  auto series_frame(SeriesID id, std::size_t index) const { return FrameView{*this, series_frame_id(id, index)}; }
  auto series_frames(SeriesID id) { return views_<FrameView>(*this, series_frame_ids(id)); }
  auto series_frames(SeriesID id) const { return views_<FrameView>(*this, series_frame_ids(id)); }
  auto series_last_frame(SeriesID id) { return FrameView{*this, series_last_frame_id(id)}; }
  auto series_last_frame(SeriesID id) const { return FrameView{*this, series_last_frame_id(id)}; }
  auto series_create_frame(SeriesID id, float64_t time) -> FrameView<Storage> { return FrameView{*this, series_create_frame_id(id, time)}; }

  // ~~ arrays ~~
  auto frame_num_arrays(FrameID id) const -> std::size_t { return need_(db_.first<std::size_t>("SELECT COUNT(*) FROM DataArrays WHERE frame_id = ?", id), "Unable to count arrays!"); }
  auto frame_array_ids(FrameID id) const -> std::vector<ArrayID> { return db_.all<ArrayID>("SELECT id FROM DataArrays WHERE frame_id = ? ORDER BY id ASC", id); }
  auto frame_find_array_id(FrameID id, std::string_view name) const -> std::optional<ArrayID> { return db_.first<ArrayID>("SELECT id FROM DataArrays WHERE frame_id = ? AND name = ?", id, name); }
  auto frame_create_array_id(FrameID id, std::string_view name) -> ArrayID {
    if (!check_frame(id)) throw Exception("Invalid frame ID!");
    if (name.empty()) throw Exception("Array name must not be empty!");
    if (frame_find_array_id(id, name)) throw Exception("Array already exists!");
    db_.run("INSERT INTO DataArrays (frame_id, name) VALUES (?, ?)", id, name);
    return ArrayID{db_.last_insert_row_id()};
  }
  void delete_array(ArrayID id) { db_.run("DELETE FROM DataArrays WHERE id = ?", id); }
  auto check_array(ArrayID id) const -> bool { return db_.first<ArrayID>("SELECT id FROM DataArrays WHERE id = ?", id).has_value(); }
  auto array_name(ArrayID id) const -> std::string { return need_(db_.first<std::string>("SELECT name FROM DataArrays WHERE id = ?", id), "Unable to get array name!"); }
  auto array_type(ArrayID id) const -> Type { return Type{need_(db_.first<std::uint32_t>("SELECT type FROM DataArrays WHERE id = ?", id), "Unable to get array type!")}; }
  auto array_size(ArrayID id) const -> std::size_t { return need_(db_.first<std::size_t>("SELECT size FROM DataArrays WHERE id = ?", id), "Unable to get array size!"); }

  auto frame_arrays(FrameID id) { return views_<ArrayView>(*this, frame_array_ids(id)); }
  auto frame_arrays(FrameID id) const { return views_<ArrayView>(*this, frame_array_ids(id)); }
  auto frame_find_array(FrameID id, std::string_view name) { return find_<Storage>(*this, id, name); }
  auto frame_find_array(FrameID id, std::string_view name) const { return find_<const Storage>(*this, id, name); }
  auto frame_create_array(FrameID id, std::string_view name) -> ArrayView<Storage> { return ArrayView{*this, frame_create_array_id(id, name)}; }

  /// Replace the contents of an array: `bytes` are packed elements of `type`.
  void array_write(ArrayID id, Type type, std::span<const std::byte> bytes) {
    if (!check_array(id)) throw Exception("Invalid array ID!");
    if (bytes.size() % type.width() != 0) throw Exception("Data size mismatch!");
    const auto packed = impl::zstd_pack(bytes);
    db_.run("UPDATE DataArrays SET type = ?, size = ?, data = ? WHERE id = ?", type.id(), bytes.size() / type.width(), std::span<const std::byte>{packed}, id);
  }
  /// Typed write: the facade's Vec / Mat are packed, so the range is its own serialisation.
  template<std::ranges::contiguous_range Range>
    requires std::ranges::sized_range<Range> && known_type_of<std::ranges::range_value_t<Range>>
  void array_write(ArrayID id, Range&& values) {
    using Val = std::ranges::range_value_t<Range>;
    static_assert(std::is_trivially_copyable_v<Val> && sizeof(Val) == type_of<Val>.width(), "elements must be packed");
    array_write(id, type_of<Val>, std::as_bytes(std::span<const Val>{std::ranges::data(values), std::ranges::size(values)}));
  }
  void array_read(ArrayID id, std::span<std::byte> bytes) const {
    if (bytes.size() != array_size(id) * array_type(id).width()) throw Exception("Data size mismatch!");
    const auto packed = db_.first<std::vector<std::byte>>("SELECT data FROM DataArrays WHERE id = ?", id);
    impl::zstd_unpack(need_(packed, "Invalid array ID!"), bytes);
  }
  auto array_read(ArrayID id) const -> std::vector<std::byte> {
    std::vector<std::byte> bytes(array_size(id) * array_type(id).width());
    array_read(id, std::span<std::byte>{bytes});
    return bytes;
  }
  template<known_type_of Val>
  auto array_read(ArrayID id) const -> std::vector<Val> {
    static_assert(std::is_trivially_copyable_v<Val> && sizeof(Val) == type_of<Val>.width(), "elements must be packed");
    if (array_type(id) != type_of<Val>) throw Exception("Type mismatch!");
    std::vector<Val> values(array_size(id));
    array_read(id, std::as_writable_bytes(std::span<Val>{values}));
    return values;
  }

private:
  impl::Db db_;
};

}  // namespace data
}  // namespace tit
