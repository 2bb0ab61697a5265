/* Thread-rank MPI shim + no-op HDF5/GSL stubs.
 *
 * TEST INFRASTRUCTURE ONLY.  Lets the reference's hot-path objects (compiled from
 * /root/reference/src where they lie, see oracle/Makefile) link and run without
 * MPI/HDF5/GSL.  Each MPI "rank" is a THREAD of the calling process:
 *
 *     pincShimInit(worldSize);            // once, before the rank threads start
 *     pincShimSetRank(r);                 // first thing in rank thread r
 *
 * Point-to-point messages are eager copies into a per-destination FIFO; matching
 * is first-in-first-out on (source, tag) with MPI_ANY_SOURCE / MPI_ANY_TAG, which
 * is what pusher.c:1015 (exchangeMigrants) relies on.  Collectives reduce in rank
 * order 0..size-1, so results are deterministic.
 */
#define _GNU_SOURCE
#include "mpi.h"
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef struct Msg_ {
	struct Msg_ *next;
	int src, tag;
	size_t bytes;
	int count;
	char *data;
} Msg;

typedef struct {
	pthread_mutex_t mu;
	pthread_cond_t cv;
	Msg *head, *tail;
} Box;

struct pincShimReq_ {
	int isRecv;
	void *buf;
	int count, type, src, tag;
};

#define SHIM_MAX_RANKS 64
#define SHIM_COLL_BYTES (1<<16)

static int g_size = 1;
static Box g_box[SHIM_MAX_RANKS];
static pthread_barrier_t g_bar;
static int g_barInit = 0;
static char g_coll[SHIM_MAX_RANKS][SHIM_COLL_BYTES];
static __thread int t_rank = 0;

static size_t typeSize(MPI_Datatype t){
	switch(t){
		case MPI_DOUBLE: return sizeof(double);
		case MPI_LONG: return sizeof(long);
		case MPI_INT: return sizeof(int);
	}
	fprintf(stderr,"mpi shim: unknown datatype %d\n",t); abort();
}

void pincShimInit(int worldSize){
	if(worldSize<1 || worldSize>SHIM_MAX_RANKS){ fprintf(stderr,"mpi shim: bad world size\n"); abort(); }
	if(g_barInit){ pthread_barrier_destroy(&g_bar); }
	for(int r=0;r<SHIM_MAX_RANKS;r++){
		pthread_mutex_init(&g_box[r].mu,NULL);
		pthread_cond_init(&g_box[r].cv,NULL);
		Msg *m = g_box[r].head;
		while(m){ Msg *n=m->next; free(m->data); free(m); m=n; }
		g_box[r].head = g_box[r].tail = NULL;
	}
	g_size = worldSize;
	pthread_barrier_init(&g_bar,NULL,worldSize);
	g_barInit = 1;
}
void pincShimSetRank(int rank){ t_rank = rank; }

int MPI_Init(int *argc, char ***argv){ (void)argc; (void)argv; return MPI_SUCCESS; }
int MPI_Finalize(void){ return MPI_SUCCESS; }
int MPI_Comm_rank(MPI_Comm c, int *rank){ (void)c; *rank = t_rank; return MPI_SUCCESS; }
int MPI_Comm_size(MPI_Comm c, int *size){ (void)c; *size = g_size; return MPI_SUCCESS; }

int MPI_Barrier(MPI_Comm c){
	(void)c;
	if(g_size>1) pthread_barrier_wait(&g_bar);
	return MPI_SUCCESS;
}

int MPI_Send(const void *buf, int count, MPI_Datatype t, int dest, int tag, MPI_Comm c){
	(void)c;
	Msg *m = malloc(sizeof(*m));
	m->next = NULL; m->src = t_rank; m->tag = tag; m->count = count;
	m->bytes = (size_t)count*typeSize(t);
	m->data = malloc(m->bytes ? m->bytes : 1);
	memcpy(m->data,buf,m->bytes);
	Box *b = &g_box[dest];
	pthread_mutex_lock(&b->mu);
	if(b->tail) b->tail->next = m; else b->head = m;
	b->tail = m;
	pthread_cond_broadcast(&b->cv);
	pthread_mutex_unlock(&b->mu);
	return MPI_SUCCESS;
}

int MPI_Recv(void *buf, int count, MPI_Datatype t, int src, int tag, MPI_Comm c, MPI_Status *st){
	(void)c;
	Box *b = &g_box[t_rank];
	pthread_mutex_lock(&b->mu);
	for(;;){
		Msg *prev = NULL, *m = b->head;
		while(m){
			if((src==MPI_ANY_SOURCE || src==m->src) && (tag==MPI_ANY_TAG || tag==m->tag)) break;
			prev = m; m = m->next;
		}
		if(m){
			if(prev) prev->next = m->next; else b->head = m->next;
			if(b->tail==m) b->tail = prev;
			pthread_mutex_unlock(&b->mu);
			size_t cap = (size_t)count*typeSize(t);
			if(m->bytes>cap){ fprintf(stderr,"mpi shim: message truncated (%zu > %zu)\n",m->bytes,cap); abort(); }
			memcpy(buf,m->data,m->bytes);
			if(st){ st->MPI_SOURCE = m->src; st->MPI_TAG = m->tag; st->MPI_ERROR = 0; st->count_ = m->count; }
			free(m->data); free(m);
			return MPI_SUCCESS;
		}
		pthread_cond_wait(&b->cv,&b->mu);
	}
}

int MPI_Isend(const void *buf, int count, MPI_Datatype t, int dest, int tag, MPI_Comm c, MPI_Request *req){
	MPI_Send(buf,count,t,dest,tag,c);          /* eager: complete on return */
	*req = MPI_REQUEST_NULL;
	return MPI_SUCCESS;
}

int MPI_Irecv(void *buf, int count, MPI_Datatype t, int src, int tag, MPI_Comm c, MPI_Request *req){
	(void)c;
	struct pincShimReq_ *r = malloc(sizeof(*r));
	r->isRecv = 1; r->buf = buf; r->count = count; r->type = t; r->src = src; r->tag = tag;
	*req = r;
	return MPI_SUCCESS;
}

int MPI_Waitall(int n, MPI_Request *reqs, MPI_Status *sts){
	(void)sts;
	for(int i=0;i<n;i++){
		if(reqs[i]==MPI_REQUEST_NULL) continue;
		struct pincShimReq_ *r = reqs[i];
		MPI_Recv(r->buf,r->count,r->type,r->src,r->tag,MPI_COMM_WORLD,MPI_STATUS_IGNORE);
		free(r);
		reqs[i] = MPI_REQUEST_NULL;
	}
	return MPI_SUCCESS;
}

int MPI_Sendrecv(const void *sbuf, int scount, MPI_Datatype st, int dest, int stag,
                 void *rbuf, int rcount, MPI_Datatype rt, int src, int rtag,
                 MPI_Comm c, MPI_Status *status){
	MPI_Send(sbuf,scount,st,dest,stag,c);
	return MPI_Recv(rbuf,rcount,rt,src,rtag,c,status);
}

static void reduceInto(void *out, int count, MPI_Datatype t, MPI_Op op){
	for(int i=0;i<count;i++){
		if(t==MPI_DOUBLE){
			double acc = ((double*)g_coll[0])[i];
			for(int r=1;r<g_size;r++){
				double v = ((double*)g_coll[r])[i];
				acc = (op==MPI_SUM) ? acc+v : (v>acc?v:acc);
			}
			((double*)out)[i] = acc;
		} else if(t==MPI_LONG){
			long acc = ((long*)g_coll[0])[i];
			for(int r=1;r<g_size;r++){
				long v = ((long*)g_coll[r])[i];
				acc = (op==MPI_SUM) ? acc+v : (v>acc?v:acc);
			}
			((long*)out)[i] = acc;
		} else {
			int acc = ((int*)g_coll[0])[i];
			for(int r=1;r<g_size;r++){
				int v = ((int*)g_coll[r])[i];
				acc = (op==MPI_SUM) ? acc+v : (v>acc?v:acc);
			}
			((int*)out)[i] = acc;
		}
	}
}

int MPI_Allreduce(const void *sbuf, void *rbuf, int count, MPI_Datatype t, MPI_Op op, MPI_Comm c){
	size_t bytes = (size_t)count*typeSize(t);
	if(bytes>SHIM_COLL_BYTES){ fprintf(stderr,"mpi shim: collective too large\n"); abort(); }
	memcpy(g_coll[t_rank], sbuf==MPI_IN_PLACE ? rbuf : sbuf, bytes);
	MPI_Barrier(c);
	reduceInto(rbuf,count,t,op);
	MPI_Barrier(c);
	return MPI_SUCCESS;
}

int MPI_Reduce(const void *sbuf, void *rbuf, int count, MPI_Datatype t, MPI_Op op, int root, MPI_Comm c){
	size_t bytes = (size_t)count*typeSize(t);
	if(bytes>SHIM_COLL_BYTES){ fprintf(stderr,"mpi shim: collective too large\n"); abort(); }
	memcpy(g_coll[t_rank], sbuf==MPI_IN_PLACE ? rbuf : sbuf, bytes);
	MPI_Barrier(c);
	if(t_rank==root) reduceInto(rbuf,count,t,op);
	MPI_Barrier(c);
	return MPI_SUCCESS;
}

int MPI_Allgather(const void *sbuf, int scount, MPI_Datatype st, void *rbuf, int rcount, MPI_Datatype rt, MPI_Comm c){
	(void)rcount; (void)rt;
	size_t bytes = (size_t)scount*typeSize(st);
	if(bytes>SHIM_COLL_BYTES){ fprintf(stderr,"mpi shim: collective too large\n"); abort(); }
	memcpy(g_coll[t_rank], sbuf, bytes);
	MPI_Barrier(c);
	for(int r=0;r<g_size;r++) memcpy((char*)rbuf + r*bytes, g_coll[r], bytes);
	MPI_Barrier(c);
	return MPI_SUCCESS;
}

/* ---- HDF5: every call collapses to this no-op (see shim/hdf5.h) ---- */
long long pincShimH5(){ return 0; }

/* ---- GSL: never drawn from by the oracle; abort loudly if it happens ---- */
typedef struct { int dummy; } gsl_rng_type_;
static gsl_rng_type_ mt = {0};
const void *gsl_rng_mt19937 = &mt;
void *gsl_rng_alloc(const void *T){ (void)T; return calloc(1,16); }
void gsl_rng_free(void *r){ free(r); }
void gsl_rng_set(const void *r, unsigned long int seed){ (void)r; (void)seed; }
double gsl_rng_uniform_pos(const void *r){ (void)r; fprintf(stderr,"gsl shim: RNG not available; inject initial conditions\n"); abort(); }
double gsl_ran_gaussian_ziggurat(const void *r, double s){ (void)r; (void)s; fprintf(stderr,"gsl shim: RNG not available; inject initial conditions\n"); abort(); }

/* ---- object.c is not compiled (does not build at this HEAD, SURVEY finding 1) ---- */
void pincShimCollision(){ }
