"""FormatAOSegmentFileName (access/appendonly/aomd.c:84-117) as cb_aocs_segfile_path states it: the column's file of a
segment file number is <relfilenode>[.<(filenum - 1) * 128 + segno>]."""
from cloudberry_b200 import capi


def test_segment_file_names():
    base = "/data/base/16384/24576"
    assert capi.aocs_segfile_path(base, 0, 1) == base                          # pseudo segno 0: no suffix
    assert capi.aocs_segfile_path(base, 1, 1) == base + ".1"
    assert capi.aocs_segfile_path(base, 127, 1) == base + ".127"
    assert capi.aocs_segfile_path(base, 0, 2) == base + ".128"                  # second column, segno 0
    assert capi.aocs_segfile_path(base, 1, 2) == base + ".129"
    assert capi.aocs_segfile_path(base, 5, 17) == base + ".%d" % (16 * 128 + 5)
    assert capi.aocs_segfile_path(base, 128, 1) is None                         # AOTupleId_MaxSegmentFileNum = 127
    assert capi.aocs_segfile_path(base, -1, 1) is None
    assert capi.aocs_segfile_path(base, 0, 0) is None                           # file numbers start at 1
    assert capi.aocs_segfile_path("x" * 5000, 1, 1) is None
