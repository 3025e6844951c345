/* integration/dropin_bodies.c -- the two function bodies a maintainer of circletile/smvp-toolkit puts in place of the
 * CPU implementations in main-cli.c to run the CSR / TJDS path on libsmvp_cuda (option B).
 * Signatures, the [INFO]/[ERROR] messages and the ownership of the returned vector are the reference's
 * (main-cli.c:325, :734); everything between the braces is new.  integration/apply_dropin.py splices the two bodies
 * into a copy of the reference's file (it locates the functions by their signatures and brace matching, so no line
 * of the reference is stored in this repository).  The markers below delimit the bodies. */

/* >>> smvp_csr_compute */
{
    /* replaces main-cli.c:336-456: qsort + CSR fill (:340-365), the timed loop (:402-420), the statistics (:428-456) */
    smvp_csr *A = NULL;
    smvp_time_stats_t st;
    int i, rc, cols = fInputRows; /* the reference sizes x by rows (main-cli.c:368) */
    double *x = (double *)malloc(sizeof(double) * (size_t)(cols > 0 ? cols : 1));
    double *y = (double *)malloc(sizeof(double) * (size_t)(fInputRows > 0 ? fInputRows : 1));
    if (!x || !y)
    {
        printf(ANSI_COLOR_RED "[ERROR]\tOut of memory.\n" ANSI_COLOR_RESET);
        exit(1);
    }
    for (i = 0; i < cols; i++)
        x[i] = 1; /* vectorInit(.., 1), main-cli.c:369 */

    printf(ANSI_COLOR_YELLOW "[INFO]\tConverting loaded content to CSR format.\n" ANSI_COLOR_RESET);
    rc = smvp_csr_build((const smvp_coo *)mmImportData, fInputRows, cols, fInputNonZeros, &A);
    if (rc != SMVP_OK)
    {
        printf(ANSI_COLOR_RED "[ERROR]\t%s %s\n" ANSI_COLOR_RESET, smvp_strerror(rc), smvp_last_cuda_error());
        exit(1);
    }
    printf(ANSI_COLOR_YELLOW "[INFO]\tCalculating %d iterations of SMVP CSR.\n" ANSI_COLOR_RESET, compiter);
    rc = smvp_csr_mult(A, x, y, compiter, csr_time->time_each, SMVP_CSR_AUTO);
    if (rc != SMVP_OK)
    {
        printf(ANSI_COLOR_RED "[ERROR]\t%s %s\n" ANSI_COLOR_RESET, smvp_strerror(rc), smvp_last_cuda_error());
        exit(1);
    }
    smvp_time_stats(csr_time->time_each, compiter, &st);
    csr_time->time_total = st.time_total;
    csr_time->time_avg = st.time_avg;
    csr_time->time_stdev = st.time_stdev;
    csr_time->time_min = st.time_min;
    csr_time->time_max = st.time_max;
    smvp_csr_free(A);
    free(x);
    return y; /* the caller prints it with generateReportText (main-cli.c:1458) */
}
/* <<< smvp_csr_compute */

/* >>> smvp_tjds_compute */
{
    /* replaces main-cli.c:755-1148: the TJDS conversion (:766-967), the timed loop (:1004-1024), the statistics
     * (:1120-1148).  SMVP_DROPIN_REF_COMPAT=1 in the environment walks only the diagonals the shipped loop walks
     * (main-cli.c:865 evaluated before :868), i.e. reproduces the reference's golden TJDS reports. */
    smvp_tjds *T = NULL;
    smvp_tjds_info_t info;
    smvp_time_stats_t st;
    int i, rc;
    const char *compat = getenv("SMVP_DROPIN_REF_COMPAT");
    double *x = (double *)malloc(sizeof(double) * (size_t)(fInputColumns > 0 ? fInputColumns : 1));
    double *y = (double *)malloc(sizeof(double) * (size_t)(fInputRows > 0 ? fInputRows : 1));
    if (!x || !y)
    {
        printf(ANSI_COLOR_RED "[ERROR]\tOut of memory.\n" ANSI_COLOR_RESET);
        exit(1);
    }
    for (i = 0; i < fInputColumns; i++)
        x[i] = 1;

    printf(ANSI_COLOR_YELLOW "[INFO]\tConverting loaded content to TJDS format.\n" ANSI_COLOR_RESET);
    rc = smvp_tjds_build((const smvp_coo *)mmImportData, fInputRows, fInputColumns, fInputNonZeros, &T);
    if (rc == SMVP_OK)
        rc = smvp_tjds_info(T, &info);
    if (rc != SMVP_OK)
    {
        printf(ANSI_COLOR_RED "[ERROR]\t%s %s\n" ANSI_COLOR_RESET, smvp_strerror(rc), smvp_last_cuda_error());
        exit(1);
    }
    printf(ANSI_COLOR_YELLOW "[INFO]\tCalculating %d iterations of SMVP TJDS.\n" ANSI_COLOR_RESET, compiter);
    rc = smvp_tjds_mult(T, x, y, compiter, tjds_time->time_each, SMVP_TJDS_ATOMIC,
                        (compat && compat[0] == '1') ? info.ref_diag_limit : 0);
    if (rc != SMVP_OK)
    {
        printf(ANSI_COLOR_RED "[ERROR]\t%s %s\n" ANSI_COLOR_RESET, smvp_strerror(rc), smvp_last_cuda_error());
        exit(1);
    }
    smvp_time_stats(tjds_time->time_each, compiter, &st);
    tjds_time->time_total = st.time_total;
    tjds_time->time_avg = st.time_avg;
    tjds_time->time_stdev = st.time_stdev;
    tjds_time->time_min = st.time_min;
    tjds_time->time_max = st.time_max;
    smvp_tjds_free(T);
    free(x);
    return y;
}
/* <<< smvp_tjds_compute */
