/*
 * pinc_h5.c - format-level HDF5 writer (see pinc_h5.h).  The structures written here follow the HDF5 File Format Specification,
 * version 1.1/2.0 subset that libhdf5 1.8 itself emits with default ("earliest") settings:
 *   superblock v0 (96 bytes, 8-byte offsets and lengths, group leaf K = 4, group internal K = 16),
 *   group = object header with a Symbol Table message -> v1 B-tree ("TREE", node type 0) -> symbol-table nodes ("SNOD", 8
 *   entries of 40 bytes) + local heap ("HEAP") holding the link names,
 *   dataset = v1 object header with Dataspace (v1), Datatype (v1, class 1 floating point: IEEE 754 binary64 little endian),
 *   Fill Value (v1, default) and Data Layout (v3, contiguous) messages; attributes = Attribute messages (v1).
 * Raw data is written as it arrives (ph5Write), all metadata when the file is closed.
 */
#define _XOPEN_SOURCE 700
#include "pinc_h5.h"
#include <fcntl.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

typedef uint64_t u64;
#define UNDEF 0xffffffffffffffffULL
#define LEAF_K 4            /* symbol-table node holds up to 2*LEAF_K entries */
#define INT_K 16            /* B-tree node holds up to 2*INT_K children */
#define DATA_START 2048     /* raw data starts here; the superblock sits at 0 */

enum { N_GROUP, N_DSET, N_XY };
typedef struct Node {
	char *name;
	int kind;
	struct Node **child; int nChild, capChild;       /* group */
	int rank; u64 dims[8]; u64 dataAddr, dataBytes;   /* dataset */
	double *rows; size_t nRows, capRows;              /* history dataset */
	u64 oh;                                           /* object header address (after emission) */
} Node;
typedef struct { char *name; double *v; int n; } Attr;
struct PH5 {
	int fd; u64 eof; int failed;
	Node *root;
	Node **dsets; int nDsets, capDsets;
	Attr *attr; int nAttr, capAttr;
};

/* ---- little-endian byte buffer --------------------------------------------------------------------------------- */
typedef struct { unsigned char *p; size_t n, cap; } Buf;
static void bput(Buf *b, const void *src, size_t n){
	if(b->n + n > b->cap){ b->cap = (b->n + n)*2 + 64; b->p = realloc(b->p, b->cap); }
	if(src) memcpy(b->p + b->n, src, n); else memset(b->p + b->n, 0, n);
	b->n += n;
}
static void b8(Buf *b, unsigned v){ unsigned char c = (unsigned char)v; bput(b, &c, 1); }
static void b16(Buf *b, unsigned v){ unsigned char c[2] = { (unsigned char)v, (unsigned char)(v >> 8) }; bput(b, c, 2); }
static void b32(Buf *b, uint32_t v){ unsigned char c[4]; for(int i = 0; i < 4; i++) c[i] = (unsigned char)(v >> (8*i)); bput(b, c, 4); }
static void b64(Buf *b, u64 v){ unsigned char c[8]; for(int i = 0; i < 8; i++) c[i] = (unsigned char)(v >> (8*i)); bput(b, c, 8); }
static void bpad8(Buf *b){ while(b->n & 7) b8(b, 0); }
static void bdbl(Buf *b, double d){ u64 v; memcpy(&v, &d, 8); b64(b, v); }        /* IEEE binary64, little endian */

static u64 align8(u64 v){ return (v + 7) & ~7ULL; }
static u64 fileAlloc(PH5 *f, u64 bytes){ u64 a = align8(f->eof); f->eof = a + bytes; return a; }
static void fileWrite(PH5 *f, u64 addr, const void *p, size_t n){
	const char *c = p;
	while(n > 0){
		ssize_t w = pwrite(f->fd, c, n, (off_t)addr);
		if(w <= 0){ f->failed = 1; return; }
		c += w; addr += (u64)w; n -= (size_t)w;
	}
}
static u64 emit(PH5 *f, Buf *b){ u64 a = fileAlloc(f, b->n); fileWrite(f, a, b->p, b->n); b->n = 0; return a; }

/* ---- the tree of links ------------------------------------------------------------------------------------------- */
static Node *newNode(const char *name, size_t len, int kind){
	Node *n = calloc(1, sizeof *n);
	n->name = malloc(len + 1); memcpy(n->name, name, len); n->name[len] = 0;
	n->kind = kind;
	return n;
}
static Node *findChild(Node *g, const char *name, size_t len){
	for(int i = 0; i < g->nChild; i++) if(strlen(g->child[i]->name) == len && !memcmp(g->child[i]->name, name, len)) return g->child[i];
	return NULL;
}
static void addChild(Node *g, Node *c){
	if(g->nChild == g->capChild){ g->capChild = g->capChild*2 + 8; g->child = realloc(g->child, (size_t)g->capChild*sizeof *g->child); }
	g->child[g->nChild++] = c;
}
/* walks "/a/b/leaf": creates the groups on the way; returns the parent group and the leaf's name (NULL if a component is not a group) */
static Node *walk(PH5 *f, const char *path, const char **leaf, size_t *leafLen, int leafIsGroup){
	Node *g = f->root;
	const char *p = path;
	while(*p == '/') p++;
	for(;;){
		const char *e = strchr(p, '/');
		size_t len = e ? (size_t)(e - p) : strlen(p);
		if(len == 0) return NULL;
		if(!e && !leafIsGroup){ *leaf = p; *leafLen = len; return g; }
		Node *c = findChild(g, p, len);
		if(!c){ c = newNode(p, len, N_GROUP); addChild(g, c); }
		if(c->kind != N_GROUP) return NULL;
		g = c;
		if(!e){ *leaf = NULL; *leafLen = 0; return g; }
		p = e + 1;
		while(*p == '/') p++;
		if(!*p){ *leaf = NULL; *leafLen = 0; return g; }
	}
}

PH5 *ph5Create(const char *path){
	int fd = open(path, O_CREAT | O_TRUNC | O_RDWR, 0644);
	if(fd < 0) return NULL;
	PH5 *f = calloc(1, sizeof *f);
	f->fd = fd; f->eof = DATA_START;
	f->root = newNode("", 0, N_GROUP);
	return f;
}
int ph5Attr(PH5 *f, const char *name, const double *value, int n){
	if(!f || n < 1) return -1;
	for(int i = 0; i < f->nAttr; i++) if(!strcmp(f->attr[i].name, name)){          /* setH5Attr overwrites (src/io.c:608-618) */
		free(f->attr[i].v); f->attr[i].v = malloc((size_t)n*sizeof(double)); memcpy(f->attr[i].v, value, (size_t)n*sizeof(double)); f->attr[i].n = n;
		return 0;
	}
	if(f->nAttr == f->capAttr){ f->capAttr = f->capAttr*2 + 4; f->attr = realloc(f->attr, (size_t)f->capAttr*sizeof *f->attr); }
	Attr *a = &f->attr[f->nAttr++];
	a->name = strdup(name); a->n = n; a->v = malloc((size_t)n*sizeof(double)); memcpy(a->v, value, (size_t)n*sizeof(double));
	return 0;
}
int ph5Group(PH5 *f, const char *path){
	const char *leaf; size_t len;
	return f && walk(f, path, &leaf, &len, 1) ? 0 : -1;
}
static int addLeaf(PH5 *f, const char *path, int kind, Node **out){
	const char *leaf; size_t len;
	Node *g = f ? walk(f, path, &leaf, &len, 0) : NULL;
	if(!g || !leaf || findChild(g, leaf, len)) return -1;
	Node *n = newNode(leaf, len, kind);
	addChild(g, n);
	if(f->nDsets == f->capDsets){ f->capDsets = f->capDsets*2 + 16; f->dsets = realloc(f->dsets, (size_t)f->capDsets*sizeof *f->dsets); }
	f->dsets[f->nDsets] = n;
	*out = n;
	return f->nDsets++;
}
int ph5Dataset(PH5 *f, const char *path, int rank, const unsigned long long *dims){
	if(rank < 1 || rank > 8) return -1;
	Node *n;
	int id = addLeaf(f, path, N_DSET, &n);
	if(id < 0) return -1;
	n->rank = rank;
	u64 count = 1;
	for(int d = 0; d < rank; d++){ n->dims[d] = dims[d]; count *= dims[d]; }
	n->dataBytes = count*8;
	n->dataAddr = count ? fileAlloc(f, n->dataBytes) : UNDEF;
	return id;
}
int ph5Write(PH5 *f, int dset, unsigned long long first, unsigned long long count, const double *data){
	if(!f || dset < 0 || dset >= f->nDsets || f->dsets[dset]->kind != N_DSET) return -1;
	Node *n = f->dsets[dset];
	if((first + count)*8 > n->dataBytes) return -1;
	/* doubles in memory are IEEE binary64; on a little-endian host (every host this library runs on) the bytes are the file's */
	fileWrite(f, n->dataAddr + first*8, data, (size_t)count*8);
	return f->failed ? -1 : 0;
}
int ph5XYCreate(PH5 *f, const char *path){
	Node *n;
	return addLeaf(f, path, N_XY, &n);
}
int ph5XYAppend(PH5 *f, int xy, double x, double y){
	if(!f || xy < 0 || xy >= f->nDsets || f->dsets[xy]->kind != N_XY) return -1;
	Node *n = f->dsets[xy];
	if(n->nRows == n->capRows){ n->capRows = n->capRows*2 + 64; n->rows = realloc(n->rows, n->capRows*2*sizeof(double)); }
	n->rows[2*n->nRows] = x; n->rows[2*n->nRows + 1] = y; n->nRows++;
	return 0;
}

/* ---- object headers (version 1) ------------------------------------------------------------------------------------ */
typedef struct { Buf b; int nMsg; } OH;
static void ohMsg(OH *h, unsigned type, unsigned flags, const Buf *data){
	size_t padded = (data->n + 7) & ~(size_t)7;
	b16(&h->b, type); b16(&h->b, (unsigned)padded); b8(&h->b, flags); b8(&h->b, 0); b8(&h->b, 0); b8(&h->b, 0);
	bput(&h->b, data->p, data->n);
	bput(&h->b, NULL, padded - data->n);
	h->nMsg++;
}
static u64 ohEmit(PH5 *f, OH *h){
	Buf o = {0};
	b8(&o, 1); b8(&o, 0); b16(&o, (unsigned)h->nMsg); b32(&o, 1); b32(&o, (uint32_t)h->b.n); b32(&o, 0);      /* 12-byte prefix + 4 bytes of alignment */
	bput(&o, h->b.p, h->b.n);
	u64 a = emit(f, &o);
	free(o.p); free(h->b.p);
	return a;
}
static void msgDataspace(Buf *b, int rank, const u64 *dims){
	b8(b, 1); b8(b, (unsigned)rank); b8(b, 0); b8(b, 0); b32(b, 0);
	for(int d = 0; d < rank; d++) b64(b, dims[d]);
}
static void msgDatatypeF64(Buf *b){
	b8(b, 0x11);                          /* version 1, class 1 (floating point) */
	b8(b, 0x20); b8(b, 0x3f); b8(b, 0);   /* little endian, zero padding, mantissa normalisation "msb implied", sign bit at 63 */
	b32(b, 8);                            /* size in bytes */
	b16(b, 0); b16(b, 64);                /* bit offset, precision */
	b8(b, 52); b8(b, 11); b8(b, 0); b8(b, 52);      /* exponent location, exponent size, mantissa location, mantissa size */
	b32(b, 1023);                         /* exponent bias */
}
static u64 emitDataset(PH5 *f, int rank, const u64 *dims, u64 addr, u64 bytes){
	OH h = {{0}, 0};
	Buf m = {0};
	msgDataspace(&m, rank, dims); ohMsg(&h, 0x0001, 0, &m); m.n = 0;
	msgDatatypeF64(&m); ohMsg(&h, 0x0003, 1, &m); m.n = 0;
	b8(&m, 1); b8(&m, 2); b8(&m, 2); b8(&m, 1); b32(&m, 0); ohMsg(&h, 0x0005, 1, &m); m.n = 0;      /* fill value v1: late allocation, write if set, defined with size 0 = the default (the bytes libhdf5 writes) */
	b8(&m, 3); b8(&m, 1); b64(&m, addr); b64(&m, bytes); ohMsg(&h, 0x0008, 0, &m); m.n = 0;           /* layout v3, contiguous */
	free(m.p);
	return ohEmit(f, &h);
}
static void msgAttr(Buf *m, const Attr *a){
	Buf dt = {0}, ds = {0};
	msgDatatypeF64(&dt);
	u64 dim = (u64)a->n; msgDataspace(&ds, 1, &dim);
	size_t nameLen = strlen(a->name) + 1;
	b8(m, 1); b8(m, 0); b16(m, (unsigned)nameLen); b16(m, (unsigned)dt.n); b16(m, (unsigned)ds.n);
	bput(m, a->name, nameLen); bpad8(m);
	bput(m, dt.p, dt.n); bpad8(m);
	bput(m, ds.p, ds.n); bpad8(m);
	for(int i = 0; i < a->n; i++) bdbl(m, a->v[i]);
	free(dt.p); free(ds.p);
}

/* ---- groups: local heap, symbol-table nodes, B-tree ---------------------------------------------------------------- */
static int byName(const void *a, const void *b){ return strcmp((*(Node *const*)a)->name, (*(Node *const*)b)->name); }
typedef struct { u64 addr, firstKey, lastKey; } TNode;

static void emitNode(PH5 *f, Node *n);
static void emitGroup(PH5 *f, Node *g, u64 *btreeOut, u64 *heapOut){
	for(int i = 0; i < g->nChild; i++) emitNode(f, g->child[i]);
	qsort(g->child, (size_t)g->nChild, sizeof *g->child, byName);
	/* local heap: the empty string at offset 0 (the key to the left of everything), then the link names, 8-byte aligned */
	Buf hd = {0};
	u64 *off = malloc(((size_t)g->nChild + 1)*sizeof *off);
	b64(&hd, 0);
	for(int i = 0; i < g->nChild; i++){ off[i] = hd.n; bput(&hd, g->child[i]->name, strlen(g->child[i]->name) + 1); bpad8(&hd); }
	u64 segSize = hd.n;
	u64 segAddr = emit(f, &hd);
	Buf b = {0};
	bput(&b, "HEAP", 4); b8(&b, 0); b8(&b, 0); b8(&b, 0); b8(&b, 0);
	b64(&b, segSize); b64(&b, 1 /* H5HL_FREE_NULL: no free block */); b64(&b, segAddr);
	*heapOut = emit(f, &b);
	/* symbol-table nodes: 2*LEAF_K entries each, names ascending */
	int nLeaf = (g->nChild + 2*LEAF_K - 1)/(2*LEAF_K);
	TNode *lvl = malloc(((size_t)nLeaf + 1)*sizeof *lvl);
	for(int s = 0; s < nLeaf; s++){
		int a = s*2*LEAF_K, e = a + 2*LEAF_K < g->nChild ? a + 2*LEAF_K : g->nChild;
		bput(&b, "SNOD", 4); b8(&b, 1); b8(&b, 0); b16(&b, (unsigned)(e - a));
		for(int i = a; i < e; i++){ b64(&b, off[i]); b64(&b, g->child[i]->oh); b32(&b, 0); b32(&b, 0); b64(&b, 0); b64(&b, 0); }
		bput(&b, NULL, (size_t)(2*LEAF_K - (e - a))*40);
		lvl[s].addr = emit(f, &b);
		lvl[s].firstKey = a ? off[a-1] : 0;      /* the key to the left of a child: the last name before it */
		lvl[s].lastKey = off[e-1];
	}
	/* B-tree levels until one node is left; a node holds up to 2*INT_K children and is stored at its full size */
	int n = nLeaf, level = 0;
	const size_t nodeBytes = 24 + (2*INT_K + 1)*8 + 2*INT_K*8;
	for(;;){
		int nNodes = n ? (n + 2*INT_K - 1)/(2*INT_K) : 1;
		TNode *up = malloc((size_t)nNodes*sizeof *up);
		u64 base = fileAlloc(f, (u64)nNodes*nodeBytes);
		for(int t = 0; t < nNodes; t++){
			int a = t*2*INT_K, e = a + 2*INT_K < n ? a + 2*INT_K : n;
			bput(&b, "TREE", 4); b8(&b, 0); b8(&b, (unsigned)level); b16(&b, (unsigned)(e - a));
			b64(&b, t ? base + (u64)(t-1)*nodeBytes : UNDEF);
			b64(&b, t + 1 < nNodes ? base + (u64)(t+1)*nodeBytes : UNDEF);
			b64(&b, e > a ? lvl[a].firstKey : 0);
			for(int i = a; i < e; i++){ b64(&b, lvl[i].addr); b64(&b, lvl[i].lastKey); }
			bput(&b, NULL, nodeBytes - b.n);
			up[t].addr = base + (u64)t*nodeBytes;
			up[t].firstKey = e > a ? lvl[a].firstKey : 0;
			up[t].lastKey = e > a ? lvl[e-1].lastKey : 0;
			fileWrite(f, up[t].addr, b.p, b.n); b.n = 0;
		}
		free(lvl); lvl = up; n = nNodes; level++;
		if(nNodes == 1) break;
	}
	*btreeOut = lvl[0].addr;
	free(lvl); free(off); free(b.p); free(hd.p);
}
static void emitNode(PH5 *f, Node *n){
	if(n->kind == N_DSET) n->oh = emitDataset(f, n->rank, n->dims, n->dataAddr, n->dataBytes);
	else if(n->kind == N_XY){
		u64 dims[2] = { n->nRows, 2 };
		u64 bytes = n->nRows*16, addr = UNDEF;
		if(bytes){ addr = fileAlloc(f, bytes); fileWrite(f, addr, n->rows, bytes); }
		n->oh = emitDataset(f, 2, dims, addr, bytes);
	} else {
		u64 bt, hp;
		emitGroup(f, n, &bt, &hp);
		OH h = {{0}, 0};
		Buf m = {0};
		b64(&m, bt); b64(&m, hp); ohMsg(&h, 0x0011, 0, &m);
		free(m.p);
		n->oh = ohEmit(f, &h);
	}
}
static void freeNode(Node *n){
	for(int i = 0; i < n->nChild; i++) freeNode(n->child[i]);
	free(n->child); free(n->rows); free(n->name); free(n);
}

int ph5Close(PH5 *f){
	if(!f) return -1;
	u64 bt, hp;
	emitGroup(f, f->root, &bt, &hp);
	OH h = {{0}, 0};
	Buf m = {0};
	b64(&m, bt); b64(&m, hp); ohMsg(&h, 0x0011, 0, &m); m.n = 0;
	for(int i = 0; i < f->nAttr; i++){ msgAttr(&m, &f->attr[i]); ohMsg(&h, 0x000c, 0, &m); m.n = 0; }
	u64 rootOH = ohEmit(f, &h);
	u64 eof = align8(f->eof);
	Buf s = {0};
	bput(&s, "\211HDF\r\n\032\n", 8);
	b8(&s, 0); b8(&s, 0); b8(&s, 0); b8(&s, 0); b8(&s, 0); b8(&s, 8); b8(&s, 8); b8(&s, 0);      /* versions, sizes of offsets and lengths */
	b16(&s, LEAF_K); b16(&s, INT_K); b32(&s, 0);
	b64(&s, 0); b64(&s, UNDEF); b64(&s, eof); b64(&s, UNDEF);      /* base, free-space info, end of file, driver info */
	b64(&s, 0); b64(&s, rootOH); b32(&s, 1); b32(&s, 0); b64(&s, bt); b64(&s, hp);      /* root symbol-table entry, cached B-tree and heap */
	fileWrite(f, 0, s.p, s.n);
	if(ftruncate(f->fd, (off_t)eof) != 0) f->failed = 1;
	int rc = f->failed ? -1 : 0;
	if(close(f->fd) != 0) rc = -1;
	free(s.p); free(m.p);
	for(int i = 0; i < f->nAttr; i++){ free(f->attr[i].name); free(f->attr[i].v); }
	free(f->attr); free(f->dsets);
	freeNode(f->root);
	free(f);
	return rc;
}
