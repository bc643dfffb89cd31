"""oracle/make_spec_ref.py -- TEST INFRASTRUCTURE: the oracle of the SPEC-CORRECT parse / write mode.

The reference has no spec-correct mode; the mode is DEFINED here as "the reference with exactly the fixes listed below", so that
its oracle is still the reference's own code.  This script reads the reference's template where it lies
(/root/reference/hevc_stream.in.c), applies the fixes as regular-expression edits (each must match the stated number of times, so a
changed reference is noticed), regenerates the C file with the reference's own generator (perl process.pl) and compiles
it with oracle/ref_harness.c into oracle/_ref/libhevcref_spec.so.  Nothing of the reference is copied into the repository: the
patched template and the generated file only exist under oracle/_ref/spec/ (git-ignored, like every other build output).

Fixes (numbers: SURVEY Appendix A; template lines of hevc_stream.in.c):
  A-1   SPS ends with rbsp_trailing_bits()                                                        (:371-377)
  A-2   a slice resolves its PPS and SPS through the id-indexed tables h->pps_table / h->sps_table  (:776-777, 924-925, 948-949)
  A-3   the derived short-term RPS variables are kept per SPS id instead of once per process        (:26-32)
  A-4   ref_pic_list_modification_flag_l1 is read / written (u1)                                   (:935)
  A-5   use_delta_flag[j] is inferred 1 when it is not present                                     (:1025-1028)
  A-6   pps_beta_offset_div2 / pps_tc_offset_div2 are present when the deblocking filter is NOT disabled  (:447)
  A-7   slice_deblocking_filter_disabled_flag inherits pps_deblocking_filter_disabled_flag         (:888)
  A-8   HRD: cpb_cnt_minus1 is present when low_delay_hrd_flag is 0, fixed_pic_rate_within_cvs_flag is inferred 1 when
        fixed_pic_rate_general_flag is 1, a sub-layer has cpb_cnt_minus1 + 1 entries; VPS: cprms_present_flag[0] is inferred 1
        (:1161-1185, 263-268)
  A-10  (generated file only) sub_layer_level_idc is u(8) on every path: the regenerated file has it, the committed one does not
Everything else (App. A 9, 11-16: scaling-list storage, slice-data handling, unsupported NAL types, missing extensions) is a
limitation of the data structures, not of the syntax walk, and stays as the reference has it.

    python oracle/make_spec_ref.py            # needs /root/reference, perl, gcc
"""
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("HEVCB_REFERENCE", "/root/reference")
OUT = os.path.join(HERE, "_ref", "spec")

# (pattern, replacement, expected matches)
EDITS = [
    # A-1: the SPS function ends with the table copy; the trailing bits go in front of it
    (r"(\n    if\( is_reading \)\n    \{\n        memcpy\(h->sps_table\[)", r"\n    structure(hevc_rbsp_trailing_bits)(b);\n\1", 1),
    # A-2: id-indexed tables
    (r"&h->pps\[sh->pic_parameter_set_id\]", r"h->pps_table[sh->pic_parameter_set_id]", 3),
    (r"&h->sps\[pps->seq_parameter_set_id\]", r"h->sps_table[pps->seq_parameter_set_id]", 3),
    # A-3: one set of derived RPS tables per SPS id; selected when an SPS announces its id and when a slice has resolved its SPS
    (r"static int NumDeltaPocs\[MAX_NUM_SHORT_TERM_REF_PICS\];\n", r"static int spec_rps_cur = 0;\nstatic int NumDeltaPocs_[32][MAX_NUM_SHORT_TERM_REF_PICS];\n#define NumDeltaPocs NumDeltaPocs_[spec_rps_cur]\n", 1),
    (r"static int NumNegativePics\[MAX_NUM_NEGATIVE_PICS\];\n", r"static int NumNegativePics_[32][MAX_NUM_NEGATIVE_PICS];\n#define NumNegativePics NumNegativePics_[spec_rps_cur]\n", 1),
    (r"static int NumPositivePics\[MAX_NUM_POSITIVE_PICS\];\n", r"static int NumPositivePics_[32][MAX_NUM_POSITIVE_PICS];\n#define NumPositivePics NumPositivePics_[spec_rps_cur]\n", 1),
    (r"static int DeltaPocS0\[MAX_NUM_REF_PICS_L0\]\[MAX_NUM_NEGATIVE_PICS\];\n", r"static int DeltaPocS0_[32][MAX_NUM_REF_PICS_L0][MAX_NUM_NEGATIVE_PICS];\n#define DeltaPocS0 DeltaPocS0_[spec_rps_cur]\n", 1),
    (r"static int UsedByCurrPicS0\[MAX_NUM_REF_PICS_L0\]\[MAX_NUM_NEGATIVE_PICS\];\n", r"static int UsedByCurrPicS0_[32][MAX_NUM_REF_PICS_L0][MAX_NUM_NEGATIVE_PICS];\n#define UsedByCurrPicS0 UsedByCurrPicS0_[spec_rps_cur]\n", 1),
    (r"static int DeltaPocS1\[MAX_NUM_REF_PICS_L1\]\[MAX_NUM_POSITIVE_PICS\];\n", r"static int DeltaPocS1_[32][MAX_NUM_REF_PICS_L1][MAX_NUM_POSITIVE_PICS];\n#define DeltaPocS1 DeltaPocS1_[spec_rps_cur]\n", 1),
    (r"static int UsedByCurrPicS1\[MAX_NUM_REF_PICS_L1\]\[MAX_NUM_POSITIVE_PICS\];\n", r"static int UsedByCurrPicS1_[32][MAX_NUM_REF_PICS_L1][MAX_NUM_POSITIVE_PICS];\n#define UsedByCurrPicS1 UsedByCurrPicS1_[spec_rps_cur]\n", 1),
    (r"(    value\( sps->sps_seq_parameter_set_id, +ue \);\n)", r"\1    spec_rps_cur = sps->sps_seq_parameter_set_id & 31;\n", 1),
    (r"(    hevc_sps_t\* sps = h->sps_table\[pps->seq_parameter_set_id\];\n\n    //set default value\n)", r"\1    spec_rps_cur = pps->seq_parameter_set_id & 31;\n", 1),
    # A-4
    (r"value\( sh->rpld\.ref_pic_list_modification_flag_l1, 1 \);", r"value( sh->rpld.ref_pic_list_modification_flag_l1, u1 );", 1),
    # A-5
    (r"(            if\( !st_ref_pic_set->used_by_curr_pic_flag\[ j \] \) \{\n                value\( st_ref_pic_set->use_delta_flag\[ j \],           u1 \);\n            \})",
     r"\1 else if( is_reading ) {\n                st_ref_pic_set->use_delta_flag[ j ] = 1;\n            }", 1),
    # A-6
    (r"if\( pps->pps_deblocking_filter_disabled_flag \) \{\n            value\( pps->pps_beta_offset_div2, se \);", r"if( !pps->pps_deblocking_filter_disabled_flag ) {\n            value( pps->pps_beta_offset_div2, se );", 1),
    # A-7
    (r"(        if\( sh->deblocking_filter_override_flag \) \{\n            value\( sh->slice_deblocking_filter_disabled_flag, u1 \);)",
     r"        if( is_reading ) { sh->slice_deblocking_filter_disabled_flag = pps->pps_deblocking_filter_disabled_flag; }\n\1", 1),
    # A-8
    (r"(        if\( !hrd->fixed_pic_rate_general_flag\[ i \] \) \{\n            value\( hrd->fixed_pic_rate_within_cvs_flag\[ i \], u1 \);\n        \})",
     r"\1 else if( is_reading ) {\n            hrd->fixed_pic_rate_within_cvs_flag[ i ] = 1;\n        }", 1),
    (r"        if\( hrd->low_delay_hrd_flag\[ i \] \) \{\n            value\( hrd->cpb_cnt_minus1\[ i \], ue \);", r"        if( !hrd->low_delay_hrd_flag[ i ] ) {\n            value( hrd->cpb_cnt_minus1[ i ], ue );", 1),
    (r"for\( int i = 0; i <= CpbCnt; i\+\+ \) \{", r"for( int i = 0; i < CpbCnt; i++ ) {", 1),
    (r"(            if \(i > 0\) \{\n                value\( vps->cprms_present_flag\[ i \],      u1 \);\n            \})",
     r"\1 else if( is_reading ) {\n                vps->cprms_present_flag[ i ] = 1;\n            }", 1),
]


def patched_template() -> str:
    text = open(os.path.join(REF, "hevc_stream.in.c")).read()
    for pat, rep, want in EDITS:
        text, got = re.subn(pat, rep, text)
        if got != want:
            raise SystemExit(f"make_spec_ref: pattern matched {got} times, expected {want}: {pat[:70]}...")
    return text


def main():
    if not os.path.isdir(REF):
        print(f"reference sources not present at {REF}: keeping prebuilt oracle/_ref/libhevcref_spec.so")
        return 0
    os.makedirs(OUT, exist_ok=True)
    tin = os.path.join(OUT, "hevc_stream.in.c")
    open(tin, "w").write(patched_template())
    with open(tin) as fi, open(os.path.join(OUT, "hevc_stream.c"), "w") as fo:
        subprocess.check_call(["perl", os.path.join(REF, "process.pl")], stdin=fi, stdout=fo)
    srcs = [os.path.join(REF, f) for f in ("hevc_nal.c", "h264_nal.c", "h264_stream.c", "h264_sei.c")]
    # -I order: the harness includes "hevc_stream.c" by name; the patched copy is found first, the reference's headers through -I REF
    cmd = ["gcc", "-O2", "-std=gnu99", "-fPIC", "-w", "-DREF_SPEC=1", "-I" + OUT, "-I" + REF, "-shared", "-o", os.path.join(HERE, "_ref", "libhevcref_spec.so"),
           os.path.join(HERE, "ref_harness.c")] + srcs + ["-lm"]
    subprocess.check_call(cmd)
    for f in ("hevc_stream.in.c", "hevc_stream.c"):  # the patched template and the generated file are intermediates: only the library stays
        os.remove(os.path.join(OUT, f))
    os.rmdir(OUT)
    print("built oracle/_ref/libhevcref_spec.so (reference template + spec fixes)")
    return 0


if __name__ == "__main__":
    sys.exit(main())
