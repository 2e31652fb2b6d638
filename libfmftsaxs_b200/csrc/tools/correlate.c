/* correlate — score the poses of an ft file against an experimental SAXS curve.
 *
 *   correlate MAPPING_PRM ATOM_PRM FTFILE RMFILE REC LIG SAXS_PROFILE L EULER_OUT OUTPUT
 *
 * Same ten arguments, side effects and output rows as the reference tool (tools/correlate.c:26-401):
 * EULER_OUT receives one "id z b1 g1 a2 b2 g2" row per ft line (three decimals), OUTPUT one row
 * "serial<TAB>ft id<TAB>chi<TAB>c1<TAB>c2" per pose whose z lies on the table 1..80 Å, in input order.
 * Where the reference spreads z steps over MPI ranks, this program is one process: the library shards
 * the z steps over the visible B200s (SXS_CUDA_DEVICES).
 */
#define _POSIX_C_SOURCE 200809L
#include <pthread.h>
#include <time.h>

#include "common.h"

#include "fftsaxs.h"
#include "index.h"

#include "mol2/atom_group.h"
#include "mol2/pdb.h"
#include "mol2/prms.h"
#include "mol2/vector.h"

/* SXS_TIMING=1: wall time of every phase on stdout (the reference prints one CPU-clock total only) */
static double wall_now(void)
{
	struct timespec ts;
	clock_gettime(CLOCK_MONOTONIC, &ts);
	return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}
static double g_phase_t0;
static void phase_done(const char *name)
{
	const double t = wall_now();
	if (getenv("SXS_TIMING") != NULL) {
		printf("[phase] %-28s %.3f s\n", name, t - g_phase_t0);
	}
	g_phase_t0 = t;
}

/* ft + rm -> Euler side file -> grid indices on its own thread: it needs the two molecule centres only, so it runs
 * beside the expansion of the molecules (SASA on host threads, CUDA start-up, K1), which it does not depend on. */
struct ft_job {
	const char *eul_path, *ft_path, *rm_path;
	struct mol_vector3 ref_lig;
	const double *zvals;
	int znum, L;
	long long nfast;
	long long *index64;
	int *ft_id, *order;
};

static void *ft_main(void *arg)
{
	struct ft_job *j = (struct ft_job *)arg;
	/* The fast route reads the ft file once with all host threads and gives the same side file and the same indices as
	 * the three passes of the reference (index.h); files it does not take go through those three passes. */
	j->nfast = sxs_ft_file_to_indices(j->eul_path, j->ft_path, j->rm_path, &j->ref_lig, j->zvals, j->znum, j->L, 0, &j->index64,
	                                  &j->ft_id, &j->order);
	if (j->nfast < 0) {
		sxs_ft_file2euler_file(j->eul_path, j->ft_path, j->rm_path, &j->ref_lig);
	}
	return NULL;
}

static void usage(void)
{
	fprintf(stderr, "Usage: correlate MAPPING_PRM ATOMPRM FTFILE RMFILE REC LIG SAXS_PROFILE L EULER_OUT OUTPUT\n");
	exit(EXIT_FAILURE);
}

static struct mol_atom_group *load_centred(const char *path, struct mol_prms *prms, int use_extrema,
                                           struct mol_vector3 *shift)
{
	struct mol_atom_group *ag = mol_read_pdb(path);
	if (ag == NULL) {
		ERROR_MSG("cannot read PDB file");
	}
	mol_atom_group_add_prms(ag, prms);
	if (use_extrema) {
		center_of_extrema(shift, ag);
	} else {
		centroid(shift, ag);
	}
	MOL_VEC_MULT_SCALAR(*shift, *shift, -1.0);
	mol_atom_group_translate(ag, shift);
	return ag;
}

int main(int argc, char *argv[])
{
	if (argc != 11) {
		usage();
	}
	char *map_path = argv[1], *prm_path = argv[2], *ft_path = argv[3], *rm_path = argv[4];
	char *rec_path = argv[5], *lig_path = argv[6], *exp_path = argv[7];
	const int L = atoi(argv[8]);
	char *eul_path = argv[9], *out_path = argv[10];

	/* translation table: 1, 2, …, 80 Å (tools/correlate.c:39-41,140-147) */
	const double z_beg = 1.0, z_end = 80.0, z_step = 1.0;
	const int znum_total = (int)round((z_end - z_beg) / z_step) + 1;
	double *zvals = (double *)calloc(znum_total, sizeof(double));
	int znum = 0;
	for (double z = z_beg; z < z_end + 0.001; z += z_step) {
		zvals[znum++] = z;
	}

	const clock_t t0 = clock();
	g_phase_t0 = wall_now();
	const int qnum = QNUM;
	double *qvals = sxs_mkarray(0.0, QMAX, qnum);

	SXS_PRINTF("Reading parameters ...\n");
	struct mol_prms *prms = mol_prms_read(prm_path);
	if (prms == NULL) {
		ERROR_MSG("cannot read atom parameter file");
	}
	SXS_PRINTF("Reading form-factors ...\n");
	struct saxs_form_factor_table *ff = default_ff_table(map_path);

	phase_done("parameter files");
	SXS_PRINTF("Reading receptor ...\n");
	struct mol_vector3 coe, com;
	struct mol_atom_group *rec = load_centred(rec_path, prms, 1, &coe);
	SXS_PRINTF("Reading ligand ...\n");
	struct mol_atom_group *lig = load_centred(lig_path, prms, 0, &com);

	/* ligand centre relative to the receptor centre in the input frames (tools/correlate.c:108-113) */
	struct mol_vector3 ref_lig;
	MOL_VEC_SUB(ref_lig, com, coe);
	MOL_VEC_MULT_SCALAR(ref_lig, ref_lig, -1.0);
	SXS_PRINTF("Converting FT and RM files into Euler coordinates ...\n");
	struct ft_job ftj = {eul_path, ft_path, rm_path, ref_lig, zvals, znum, L, -1, NULL, NULL, NULL};
	pthread_t ft_thread;
	const int ft_threaded = pthread_create(&ft_thread, NULL, ft_main, &ftj) == 0;
	if (!ft_threaded) {
		ft_main(&ftj);
	}

	struct sxs_spf_full *A = atom_grp2spf(rec, ff, qvals, qnum, L, 1);
	struct sxs_spf_full *B = atom_grp2spf(lig, ff, qvals, qnum, L, 1);
	phase_done("PDB + SASA + expansion (K1)");
	if (ft_threaded) {
		pthread_join(ft_thread, NULL);
	}
	long long *index64 = ftj.index64;
	int *ft_id = ftj.ft_id, *order = ftj.order;
	const long long nfast = ftj.nfast;

	phase_done("ft + rm -> Euler file + indices (rest)");
	SXS_PRINTF("Reading experiment ...\n");
	struct sxs_profile *exp_profile = sxs_profile_read(exp_path);
	if (exp_profile == NULL) {
		ERROR_MSG("cannot read experimental profile");
	}
	const double mean_radius = (A->rm * rec->natoms + B->rm * lig->natoms) / (rec->natoms + lig->natoms);
	struct sxs_opt_params *params = sxs_opt_params_create(exp_profile, qvals, qnum, mean_radius);
	free(exp_profile->qvals);
	sxs_profile_free(exp_profile);
	mol_prms_free(prms);
	mol_atom_group_free(rec);
	mol_atom_group_free(lig);

	/* Euler rows -> grid indices; rows off the z table are dropped, serial numbers count every line */
	size_t n = 0;
	int *index = NULL;
	/* The reference packs the six digits into an `int` (tools/correlate.c:225-240), which overflows from L = 20 on with
	 * the 80-step z table ((L+1)^2 (2L+1)^3 * 80 > 2^31): its rows are then wrong.  Where 32 bits suffice this tool
	 * does exactly what the reference does; beyond that it keeps the 64-bit indices of the fast route. */
	const double index_space = (double)znum * (L + 1) * (L + 1) * (2.0 * L + 1) * (2.0 * L + 1) * (2.0 * L + 1);
	const int wide = nfast >= 0 && index_space > 2147483647.0;
	SXS_PRINTF("Reading Euler coordinates ...\n");
	if (wide) {
		n = (size_t)nfast;
	} else if (nfast >= 0) {
		n = (size_t)nfast;
		index = (int *)malloc((n ? n : 1) * sizeof(int));
		CHECK_PTR(index);
		for (size_t i = 0; i < n; i++) {
			index[i] = (int)index64[i];
		}
		free(index64);
		index64 = NULL;
	} else {
		FILE *ef = fopen(eul_path, "r");
		if (ef == NULL) {
			ERROR_MSG("cannot reopen Euler file");
		}
		size_t cap = 1 << 16;
		index = (int *)malloc(cap * sizeof(int));
		ft_id = (int *)malloc(cap * sizeof(int));
		order = (int *)malloc(cap * sizeof(int));
		struct sxs_euler e;
		int id, line = 0;
		while (fscanf(ef, "%d %lf %lf %lf %lf %lf %lf", &id, &e.z, &e.b1, &e.g1, &e.a2, &e.b2, &e.g2) != EOF) {
			for (int j = 0; j < znum; j++) {
				if (zvals[j] > e.z - 0.001 && zvals[j] < e.z + 0.001) {
					if (n == cap) {
						cap *= 2;
						index = (int *)realloc(index, cap * sizeof(int));
						ft_id = (int *)realloc(ft_id, cap * sizeof(int));
						order = (int *)realloc(order, cap * sizeof(int));
						CHECK_PTR(index); CHECK_PTR(ft_id); CHECK_PTR(order);
					}
					index[n] = sxs_euler_to_index(&e, j, L);
					ft_id[n] = id;
					order[n] = line;
					n++;
					/* the reference reflects a2/g2 in place inside this loop, so a second matching z would see the
					 * reflected angles; with the 1 Å table at most one z matches */
				}
			}
			line++;
		}
		fclose(ef);
	}

	phase_done("experiment + index arrays");
	double *score = (double *)calloc(n ? n : 1, sizeof(double));
	double *c1 = (double *)calloc(n ? n : 1, sizeof(double));
	double *c2 = (double *)calloc(n ? n : 1, sizeof(double));
	SXS_PRINTF("\nCORRELATION STARTED\n\n");
	if (wide) {
		sxs_compute_saxs_scores64(score, c1, c2, index64, (long long)n, A, B, params, qvals, qnum, zvals, znum, L, 1);
	} else {
		sxs_compute_saxs_scores(score, c1, c2, index, (int)n, A, B, params, qvals, qnum, zvals, znum, L, 1);
	}

	phase_done("scoring (K2-K4 + copies)");
	printf("\nTime passed: %.3f\n", (double)(clock() - t0) / CLOCKS_PER_SEC);
	printf("Writing results to %s\n", out_path);
	sxs_write_score_rows(out_path, (long long)n, order, ft_id, score, c1, c2, 0);
	phase_done("output rows");
	printf("\nCorrelation finished\n");

	free(score); free(c1); free(c2); free(index); free(index64); free(ft_id); free(order);
	sxs_opt_params_free(params);
	sxs_spf_full_free(A);
	sxs_spf_full_free(B);
	free(zvals);
	free(qvals);
	return EXIT_SUCCESS;
}
