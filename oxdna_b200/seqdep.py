"""Published sequence-dependent parameter tables of oxDNA2 and oxRNA (the contents of the reference's
oxDNA2_sequence_dependent_parameters.txt and rna_sequence_dependent_parameters.txt), usable as `seq_dep_file` values of
`Simulation` and writable as files for the reference's own binaries.  Base order in the keys: A, G, C, T (U for RNA)."""

DNA2_SEQ_DEP = {
    "STCK_FACT_EPS": 0.18,
    "STCK_G_C": 1.69339, "STCK_C_G": 1.74669, "STCK_G_G": 1.61295, "STCK_C_C": 1.61295, "STCK_G_A": 1.59887, "STCK_T_C": 1.59887,
    "STCK_A_G": 1.61898, "STCK_C_T": 1.61898, "STCK_T_G": 1.66322, "STCK_C_A": 1.66322, "STCK_G_T": 1.68032, "STCK_A_C": 1.68032,
    "STCK_A_T": 1.56166, "STCK_T_A": 1.64311, "STCK_A_A": 1.84642, "STCK_T_T": 1.58952,
    "HYDR_A_T": 0.88537, "HYDR_T_A": 0.88537, "HYDR_C_G": 1.23238, "HYDR_G_C": 1.23238,
}

RNA_SEQ_DEP = {
    "HYDR_A_T": 0.820419, "HYDR_C_G": 1.06444, "HYDR_G_T": 0.510558,
    "STCK_G_C": 1.27562, "STCK_C_G": 1.60302, "STCK_G_G": 1.49422, "STCK_C_C": 1.47301, "STCK_G_A": 1.62114, "STCK_T_C": 1.16724,
    "STCK_A_G": 1.39374, "STCK_C_T": 1.47145, "STCK_T_G": 1.28576, "STCK_C_A": 1.58294, "STCK_G_T": 1.57119, "STCK_A_C": 1.21041,
    "STCK_A_T": 1.38529, "STCK_T_A": 1.24573, "STCK_A_A": 1.31585, "STCK_T_T": 1.17518,
    **{f"CROSS_{a}_{b}": 59.9626 for a in "AGCT" for b in "AGCT"},
    "ST_T_DEP": 1.97561,
}


def write_file(path, table):
    with open(path, "w") as f:
        for k, v in table.items():
            f.write(f"{k} = {v!r}\n")
    return path
