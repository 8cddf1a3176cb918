"""Golden facts about the reference's own `.ttdb` fixture
(/root/reference/tests/_data/particles.ttdb, written by the reference's
Storage / ZSTD stream compressor) for tests/test_ttdb.py.

Run in the build container (the reference tree does not travel to the GPU box):
    python tests/golden/make_ttdb_golden.py
Writes tests/golden/ttdb_fixture.json: per array its name, type id, size and the
SHA-1 of the *decompressed* bytes (decompressed here with the zstd CLI-independent
`pyarrow` codec, i.e. not with the code under test), plus two of the reference's
compressed blobs verbatim (base64, a few kB) so that the readers are exercised on
frames produced by the reference's streaming compressor.
"""
import base64
import hashlib
import json
import os
import sqlite3

import pyarrow as pa

SRC = "/root/reference/tests/_data/particles.ttdb"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ttdb_fixture.json")
WIDTH = {1: 1, 2: 1, 3: 2, 4: 2, 5: 4, 6: 4, 7: 8, 8: 8, 9: 4, 10: 8}  # by kind + 1


def main():
    db = sqlite3.connect(f"file:{SRC}?mode=ro", uri=True)
    codec = pa.Codec("zstd")
    out = {"source": "tests/_data/particles.ttdb", "tables": sorted(n for (n,) in db.execute("SELECT name FROM sqlite_master WHERE type = 'table' AND name NOT LIKE 'sqlite_%'")),
           "max_series": db.execute("SELECT max_series FROM Settings").fetchone()[0], "series": []}
    for sid, sname in db.execute("SELECT id, name FROM DataSeries ORDER BY id"):
        frames = []
        for fid, time in db.execute("SELECT id, time FROM DataFrames WHERE series_id = ? ORDER BY id", (sid,)):
            arrays = []
            for aid, name, tid, size, blob in db.execute("SELECT id, name, type, size, data FROM DataArrays WHERE frame_id = ? ORDER BY id", (fid,)):
                rank, dim = (tid >> 8) & 0xFF, (tid >> 16) & 0xFF
                nbytes = size * WIDTH[tid & 0xFF] * dim ** rank
                raw = codec.decompress(blob, decompressed_size=nbytes).to_pybytes()
                assert len(raw) == nbytes
                rec = {"id": aid, "name": name, "type": tid, "size": size, "nbytes": nbytes, "sha1": hashlib.sha1(raw).hexdigest()}
                if (fid, name) in ((2, "FS"), (1, "rho")):
                    rec["blob_b64"] = base64.b64encode(blob).decode()
                arrays.append(rec)
            frames.append({"id": fid, "time": time, "arrays": arrays})
        out["series"].append({"id": sid, "name": sname or "", "frames": frames})
    with open(OUT, "w") as f:
        json.dump(out, f, indent=1)
    print(OUT, os.path.getsize(OUT))


if __name__ == "__main__":
    main()
