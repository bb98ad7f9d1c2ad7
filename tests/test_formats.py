"""Persistence formats of the reference node restated without PCL / Boost — pure host functions of
libm3dreg.so, no GPU needed:
  * binary PCD of the 40-byte PointXYZIRNLRGB as pcl::io::savePCDFileBinary writes a typed cloud (src/gpu6DSLAM.cpp:41,88),
  * the XML pose model of class data_model (src/data_model.cpp) as boost::property_tree::write_xml lays it out.
The expected bytes below are written out by hand from the formats' definitions (PCL pcd_io.hpp generateHeader<PointT> /
writeBinary<PointT>; property_tree xml_parser write with '\\t' x 1 indentation), not produced by the code under test."""
import struct

import numpy as np


def _cloud(synth, n=37):
    c = synth.random_cloud(n, seed=5)
    c["ring"] = np.arange(n) % 32
    c["intensity"] = np.linspace(0, 255, n, dtype=np.float32)
    c["rgb"] = np.arange(n, dtype=np.float32) * 0.5
    return c


PCD_HEADER = ("# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\nFIELDS x y z intensity ring normal_x normal_y normal_z label rgb\n"
              "SIZE 4 4 4 4 2 4 4 4 4 4\nTYPE F F F F U F F F I F\nCOUNT 1 1 1 1 1 1 1 1 1 1\nWIDTH {n}\nHEIGHT 1\nVIEWPOINT 0 0 0 1 0 0 0\n"
              "POINTS {n}\nDATA binary\n")


def _packed(c):
    return b"".join(struct.pack("<ffffHfffif", float(p["x"]), float(p["y"]), float(p["z"]), float(p["intensity"]), int(p["ring"]),
                                float(p["normal_x"]), float(p["normal_y"]), float(p["normal_z"]), int(p["label"]), float(p["rgb"])) for p in c)


def test_pcd_binary_layout_and_round_trip(pkg, synth, tmp_path):
    c = _cloud(synth)
    path = tmp_path / "scan.pcd"
    pkg.pcd_write_binary(path, c)
    raw = path.read_bytes()
    hdr = PCD_HEADER.format(n=len(c)).encode()
    assert raw[:len(hdr)] == hdr
    assert raw[len(hdr):] == _packed(c)                      # 38 packed bytes per point, the registered fields only
    back = pkg.pcd_read(path)
    for f in c.dtype.names:
        assert np.array_equal(back[f].view(np.uint32 if back[f].dtype.itemsize == 4 else back[f].dtype),
                              c[f].view(np.uint32 if c[f].dtype.itemsize == 4 else c[f].dtype)), f
    # empty cloud
    pkg.pcd_write_binary(tmp_path / "empty.pcd", c[:0])
    assert len(pkg.pcd_read(tmp_path / "empty.pcd")) == 0


def test_pcd_reader_other_layouts(pkg, synth, tmp_path):
    c = _cloud(synth, 11)
    # (a) the padded 40-byte layout the PCLPointCloud2 writer produces (`_` stands for the struct's padding), fields reordered
    hdr = ("# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\nFIELDS x y z intensity ring _ normal_x normal_y normal_z label rgb\n"
           "SIZE 4 4 4 4 2 1 4 4 4 4 4\nTYPE F F F F U U F F F I F\nCOUNT 1 1 1 1 1 2 1 1 1 1 1\nWIDTH 11\nHEIGHT 1\nVIEWPOINT 0 0 0 1 0 0 0\nPOINTS 11\nDATA binary\n")
    (tmp_path / "padded.pcd").write_bytes(hdr.encode() + c.tobytes())
    back = pkg.pcd_read(tmp_path / "padded.pcd")
    assert all(np.array_equal(back[f], c[f]) for f in c.dtype.names)
    # (b) an ASCII file of a plain XYZI cloud (another PCL point type): missing fields read as zero
    lines = ["VERSION .7", "FIELDS x y z intensity", "SIZE 4 4 4 4", "TYPE F F F F", "COUNT 1 1 1 1", "WIDTH 3", "HEIGHT 1", "POINTS 3", "DATA ascii",
             "1.5 -2 0.25 7", "0 0 0 0", "-1e-3 4 5 255"]
    (tmp_path / "ascii.pcd").write_text("\n".join(lines) + "\n")
    back = pkg.pcd_read(tmp_path / "ascii.pcd")
    assert back["x"].tolist() == [1.5, 0.0, np.float32(-1e-3)] and back["intensity"].tolist() == [7.0, 0.0, 255.0]
    assert not back["normal_x"].any() and not back["label"].any()
    # (c) rgb stored as U32 (newer PCL writers): same bits
    hdr = ("VERSION 0.7\nFIELDS x y z rgb\nSIZE 4 4 4 4\nTYPE F F F U\nCOUNT 1 1 1 1\nWIDTH 1\nHEIGHT 1\nPOINTS 1\nDATA binary\n")
    (tmp_path / "rgbu.pcd").write_bytes(hdr.encode() + struct.pack("<fffI", 1.0, 2.0, 3.0, 0x00FF8040))
    back = pkg.pcd_read(tmp_path / "rgbu.pcd")
    assert back["rgb"].view(np.uint32)[0] == 0x00FF8040
    # (d) errors: missing file, no xyz
    import pytest
    with pytest.raises(pkg.M3dRegError):
        pkg.pcd_read(tmp_path / "nope.pcd")
    (tmp_path / "bad.pcd").write_text("VERSION .7\nFIELDS a b\nSIZE 4 4\nTYPE F F\nCOUNT 1 1\nWIDTH 1\nHEIGHT 1\nPOINTS 1\nDATA ascii\n1 2\n")
    with pytest.raises(pkg.M3dRegError):
        pkg.pcd_read(tmp_path / "bad.pcd")


XML_EXPECTED = """<?xml version="1.0" encoding="utf-8"?>
<Model>
\t<Algorithms>
\t\t<name>registration: semantic point to point</name>
\t</Algorithms>
\t<DatasetPath>processedData</DatasetPath>
\t<Transformations>
\t\t<scan_A>
\t\t\t<Affine>
\t\t\t\t<Type>matrix4f</Type>
\t\t\t\t<Data>1 0 0 0 0 0.5 -0.25 0 0 0.25 0.5 0 1.5 -2 1e-05 1 </Data>
\t\t\t</Affine>
\t\t\t<cloudname>scan_A.pcd</cloudname>
\t\t</scan_A>
\t\t<scan_B>
\t\t\t<cloudname>scan_B.pcd</cloudname>
\t\t\t<Affine>
\t\t\t\t<Type>matrix4f</Type>
\t\t\t\t<Data>1 0 0 0 0 1 0 0 0 0 1 0 123457 0.333333 3 1 </Data>
\t\t\t</Affine>
\t\t</scan_B>
\t</Transformations>
</Model>
"""


def test_xml_model_layout_and_round_trip(pkg, tmp_path):
    m = pkg.Model()
    m.set_algorithm_name("registration: semantic point to point")
    m.set_dataset_path("processedData")
    a = np.array([[1, 0, 0, 1.5], [0, 0.5, 0.25, -2], [0, -0.25, 0.5, 1e-5], [0, 0, 0, 1]], dtype=np.float32)
    b = np.eye(4, dtype=np.float32); b[:3, 3] = [123456.7, 1.0 / 3.0, 3.0]
    m.set_affine("scan_A", a); m.set_cloud_name("scan_A", "scan_A.pcd")
    m.set_cloud_name("scan_B", "scan_B.pcd"); m.set_affine("scan_B", b)          # insertion order is kept, as a ptree does
    path = tmp_path / "sub" / "registeredData_x.xml"
    path.parent.mkdir()
    m.save(path)
    assert path.read_text() == XML_EXPECTED                # column-major data, operator<<(float): six significant digits
    r = pkg.Model()
    assert r.load(path)
    assert r.scan_ids() == ["scan_A", "scan_B"]
    assert np.array_equal(r.affine("scan_A"), a)
    assert np.allclose(r.affine("scan_B"), b, rtol=1e-5) and r.affine("scan_B")[0, 3] == np.float32(123457.0)     # the format is lossy by definition
    assert r.cloud_name("scan_B") == "scan_B.pcd" and r.dataset_path() == "processedData"
    assert r.full_cloud_path("scan_A") == str(path.parent / "processedData" / "scan_A.pcd")
    assert r.affine("scan_C") is None and r.cloud_name("scan_C") is None
    assert not pkg.Model().load(tmp_path / "missing.xml")
    # overwriting a value keeps the node's place
    m.set_affine("scan_A", np.eye(4, dtype=np.float32))
    m.save(path)
    assert pkg.Model().load(path) and path.read_text().index("scan_A") < path.read_text().index("scan_B")


def test_xml_reader_accepts_quaternion_type_and_escapes(pkg, tmp_path):
    txt = """<?xml version="1.0" encoding="utf-8"?>
<!-- hand-written -->
<Model><DatasetPath>raw &amp; more</DatasetPath><Transformations>
<s0><Affine><Type>Vector3f_Quaternionf</Type><Data>1 2 3 0 0 0.70710678 0.70710678</Data></Affine><cloudname>a&lt;b&gt;.pcd</cloudname></s0>
<s1><Affine><Type>matrix4f</Type><Data>1 0 0 0  0 1 0 0  0 0 1 0  4 5 6 1</Data></Affine></s1>
<s2/></Transformations></Model>"""
    p = tmp_path / "m.xml"
    p.write_text(txt)
    m = pkg.Model()
    assert m.load(p)
    assert m.scan_ids() == ["s0", "s1", "s2"] and m.dataset_path() == "raw & more" and m.cloud_name("s0") == "a<b>.pcd"
    q = m.affine("s0")                                        # 90 degrees about Z, origin (1, 2, 3): data_model.cpp:75-87
    assert np.allclose(q, [[0, -1, 0, 1], [1, 0, 0, 2], [0, 0, 1, 3], [0, 0, 0, 1]], atol=1e-6)
    assert np.array_equal(m.affine("s1")[:3, 3], [4, 5, 6]) and m.affine("s2") is None
    (tmp_path / "broken.xml").write_text("<Model><a></Model>")
    assert not pkg.Model().load(tmp_path / "broken.xml")
