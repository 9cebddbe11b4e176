"""Layout of the reference's time-series files (SaveTimeSrs, src/file_manip.f:741-830), written by the host side."""
import numpy as np

from wolfd2_b200.timeseries import TimeSeriesWriter


def test_header_and_rows_cold_flow(tmp_path):
    w = TimeSeriesWriter(str(tmp_path / "run"), [5, 12], [7, 3])
    assert [f.split("/")[-1] for f in w.filenames()] == ["run001.ts", "run002.ts"]
    rec = np.arange(16, dtype=np.float64).reshape(2, 8) * 0.125 - 0.5
    with w:
        w.write(0.25, rec)
        w.write(0.5, -rec)
    lines = open(w.filenames()[1]).read().split("\n")
    assert lines[0] == "#" and lines[3] == "#"
    assert lines[1] == "# Time-series No.    2"                       # '(a,i4)'
    assert lines[2] == "# Location:   12   3"                         # '(a,2i4)'
    assert lines[4] == "#    Time" + "             u" + "             v" + "             P"      # '(a,8a14)'
    # '(17(e14.6))': time, u, v, p of point 2 = rec[1][0:3] = 0.5, 0.625, 0.75
    assert lines[5] == "  0.250000E+00" + "  0.500000E+00" + "  0.625000E+00" + "  0.750000E+00"
    assert lines[6] == "  0.500000E+00" + " -0.500000E+00" + " -0.625000E+00" + " -0.750000E+00"
    assert lines[7] == "" and len(lines) == 8


def test_columns_follow_the_flags(tmp_path):
    rec = np.arange(8, dtype=np.float64).reshape(1, 8) + 1.0
    for thermal, small, want in ((True, False, [1, 2, 3, 4]), (False, True, [1, 2, 3, 5, 6, 7]),
                                 (True, True, [1, 2, 3, 4, 5, 6, 7, 8])):
        w = TimeSeriesWriter(str(tmp_path / f"s{int(thermal)}{int(small)}_"), [2], [2], thermal=thermal, smallscale=small)
        with w:
            w.write(1.0, rec)
        lines = open(w.filenames()[0]).read().split("\n")
        names = lines[4][9:].split()
        assert len(names) == len(want)
        assert names[:3] == ["u", "v", "P"] and (("T" in names) == thermal) and (("u*" in names) == small)
        vals = [float(lines[5][14 * (k + 1):14 * (k + 2)]) for k in range(len(want))]
        assert vals == [float(v) for v in want]
