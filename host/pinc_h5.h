/*
 * pinc_h5.h - a format-level HDF5 writer for the output PINC's host writes, without libhdf5.
 *
 * libhdf5 is not in this image, so the C host cannot call H5Fcreate/H5Dwrite as the reference does (src/io.c:566-739,
 * src/grid.c:1161-1270, src/population.c:497-651).  This file writes the subset of the HDF5 file format those calls produce:
 * superblock version 0, old-style groups (local heap + version-1 B-tree + symbol-table nodes), version-1 object headers,
 * contiguous datasets of IEEE little-endian doubles and double-valued attributes on the root group.  Host code, no device
 * work: the data it writes is what pincSyncGridToHost / pincSyncPopToHost stage.
 *
 * What is the same as in the reference's files: file names (<prefix><name>.grid.h5 / .pop.h5 / .xy.h5), group and dataset
 * names ("/n=%.1f", "/pos/specie %i/n=%.1f", "/energy/..."), datatype (H5T_IEEE_F64LE), dataspace extents (dimensions
 * reversed, (z,y,x,component), true nodes only), attribute names and values.  What differs: the history datasets are written
 * as contiguous (N,2) arrays when the file is closed (the reference makes them chunked and extendible, {1,2} chunks); one
 * process writes a file (no MPI-IO).  Verified with an independent reader that is itself pinned to a file written by the real
 * library (tests/h5mini.py, tests/test_h5_writer.py).
 */
#ifndef PINC_H5_H
#define PINC_H5_H
#ifdef __cplusplus
extern "C" {
#endif

typedef struct PH5 PH5;

/* creates (truncates) the file; NULL on failure */
PH5 *ph5Create(const char *path);
/* attribute of n doubles on the root group, as setH5Attr(file, name, value, size) (src/io.c:606); 0 on success */
int ph5Attr(PH5 *f, const char *name, const double *value, int n);
/* group path "/a/b" with its parents, as createH5Group / H5Gcreate; 0 on success */
int ph5Group(PH5 *f, const char *path);
/* dataset of doubles at `path` (parents are created), extents dims[0..rank) slowest first as H5Screate_simple takes them;
 * the raw data area is reserved in the file.  Returns a dataset handle >= 0, or -1 */
int ph5Dataset(PH5 *f, const char *path, int rank, const unsigned long long *dims);
/* `count` consecutive elements of dataset `dset`, starting at element `first` (row-major position), from `data` */
int ph5Write(PH5 *f, int dset, unsigned long long first, unsigned long long count, const double *data);
/* history dataset as xyCreateDataset / xyWrite (src/io.c:657-733): rows (x, y) are collected and written as an (N,2) dataset on close */
int ph5XYCreate(PH5 *f, const char *path);
int ph5XYAppend(PH5 *f, int xy, double x, double y);
/* writes groups, object headers and the superblock; the file is a valid HDF5 file only after this; 0 on success */
int ph5Close(PH5 *f);

#ifdef __cplusplus
}
#endif
#endif
