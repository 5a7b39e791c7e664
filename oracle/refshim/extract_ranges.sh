#!/bin/bash
# oracle/refshim/extract_ranges.sh REF OUT -- TEST INFRASTRUCTURE ONLY (build step of oracle/_ref/libdcmt_ref.so).
#
# The stereo and evaluation functions of the reference live inside whole-program files (main_sl.cpp, main_lc.cpp,
# main.cpp: PCL viewers, file IO, main()), which cannot be compiled here.  This step streams the line ranges that hold
# just those functions out of the reference tree into a scratch include file OUT (under /tmp, deleted by the Makefile
# after the compile; never written into the repository), after checking that each range still starts at the
# function it is supposed to hold.
set -euo pipefail
REF="$1"; OUT="$2"; OUTDIR="$(dirname "$OUT")"
SL="$REF/src/DC_stereo_lidar/main_sl.cpp"; LC="$REF/src/DC_lidar_camera/main_lc.cpp"; LO="$REF/src/DC_lidar_only/main.cpp"
anchor() {  # file line expected-prefix
    local got; got="$(sed -n "${2}p" "$1")"
    case "$got" in "$3"*) ;; *) echo "extract_ranges: $1:$2 does not start with '$3' (reference changed?)" >&2; exit 1;; esac
}
anchor "$SL" 23 "struct EntryType"
anchor "$SL" 715 "void calculateMeasuementDerivatives"
anchor "$SL" 747 "bool calculateObservationDerivatives"
anchor "$SL" 804 "void optimize_IG"
anchor "$SL" 846 "void get_initial_disparity"
anchor "$SL" 863 "void retrieve_optimized_depth"
anchor "$SL" 1031 "void evaluate_performances"
anchor "$LC" 85 "void evaluate_performance"
anchor "$LO" 16 "void evaluate_performance"
UT="$REF/src/DC_lidar_only/utils.cpp"
anchor "$SL" 474 "    pcl::PointCloud<pcl::PointXYZ>::Ptr transformedCloud"
anchor "$SL" 478 "    for (size_t i = 0; i < cloud->size(); ++i){"
anchor "$SL" 523 "    cv::normalize(projected_depths, normalized_depths, 0, 80, cv::NORM_MINMAX);"
anchor "$UT" 15 "void read_M"
anchor "$UT" 39 "void write_M"
{
    echo "// generated from $REF by oracle/refshim/extract_ranges.sh -- scratch file, do not keep"
    echo "#line 23 \"$SL\"";   sed -n '23,26p' "$SL"
    echo "#line 715 \"$SL\"";  sed -n '715,885p' "$SL"
    echo "#line 1031 \"$SL\""; sed -n '1031,1061p' "$SL"
    echo "#line 85 \"$LC\"";   sed -n '85,116p' "$LC"
    echo "#line 16 \"$LO\"";   sed -n '16,34p' "$LO"
} > "$OUT"
# the LiDAR projection loops + cv::normalize of withSuperPixels (main_sl.cpp:474-523): a function BODY, included inside a
# wrapper that declares the variables it uses (refshim_front.cpp)
{
    echo "// generated from $REF by oracle/refshim/extract_ranges.sh -- scratch file, do not keep"
    echo "#line 474 \"$SL\""; sed -n '474,523p' "$SL"
} > "$OUTDIR/ref_project.inc"
# read_M / write_M (utils.cpp:15-58)
{
    echo "// generated from $REF by oracle/refshim/extract_ranges.sh -- scratch file, do not keep"
    echo "#line 15 \"$UT\""; sed -n '15,58p' "$UT"
} > "$OUTDIR/ref_utils.inc"
