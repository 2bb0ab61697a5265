/*
 * pinc_oracle.c — CPU restatement of PINC's per-timestep PIC loop (see pinc_oracle.h).
 *
 * TEST INFRASTRUCTURE ONLY: never on the product path.  Parity status: PINNED against the
 * reference's known-answer vectors and against the reference's own sources compiled in
 * place (oracle/_ref), see tests/test_oracle_*.py.
 *
 * Written from the algorithm, not from the text of the reference: loops run over explicit
 * (j,k,l) node indices; the order of floating point operations inside every expression
 * follows the cited reference line so results are bit-identical where the reference is
 * deterministic.  Compile with -ffp-contract=off (the reference is built without FMA).
 */
#include "pinc_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define IDX(j,k,l,sx,sy) ((long)(j) + (long)(sx)*((long)(k) + (long)(sy)*(long)(l)))

/* ------------------------------------------------------------------------------------------
 * Particles
 * ---------------------------------------------------------------------------------------- */

/* src/pusher.c:86-119 (quirk Q4: the collision loop is dead code; what runs is pos += vel). */
void orc_move(double *pos, const double *vel, int nSpecies, const long *iStart, const long *iStop){
	for(int s=0;s<nSpecies;s++)
		for(long p=3*iStart[s]; p<3*iStop[s]; p++) pos[p] += vel[p];
}

/* src/grid.c:668-673 */
void orc_gmul(double *val, long n, double num){
	for(long p=0;p<n;p++) val[p] *= num;
}

/* src/pusher.c:1089-1122: trilinear gather of the three field components. */
static void interp3d1(double *dv, const double *pos, const double *E, long sx, long sy){
	int j = (int)pos[0], k = (int)pos[1], l = (int)pos[2];
	double x = pos[0]-j, y = pos[1]-k, z = pos[2]-l;
	double xc = 1-x, yc = 1-y, zc = 1-z;
	long p000 = 3*IDX(j,k,l,sx,sy);
	long p100 = p000+3, p010 = p000+3*sx, p110 = p010+3;
	long p001 = p000+3*sx*sy, p101 = p001+3, p011 = p001+3*sx, p111 = p011+3;
	for(int v=0;v<3;v++)
		dv[v] = zc*( yc*(xc*E[p000+v]+x*E[p100+v]) + y*(xc*E[p010+v]+x*E[p110+v]) )
		      + z *( yc*(xc*E[p001+v]+x*E[p101+v]) + y*(xc*E[p011+v]+x*E[p111+v]) );
}

/* src/pusher.c:147-176 (kinEnergy==NULL) and :178-214 (with the mid-step energy).
 * The whole E grid is rescaled by q/m before and m/q after each species (quirk Q2). */
void orc_acc3d1(double *pos, double *vel, int nSpecies, const long *iStart, const long *iStop,
                const double *charge, const double *mass, double *E, const int *size, double *kinEnergy){
	long sx = size[0], sy = size[1], n = 3L*size[0]*size[1]*size[2];
	for(int s=0;s<nSpecies;s++){
		orc_gmul(E,n,charge[s]/mass[s]);
		double acc = 0;
		for(long p=3*iStart[s]; p<3*iStop[s]; p+=3){
			double dv[3];
			interp3d1(dv,&pos[p],E,sx,sy);
			double v2 = 0;
			for(int d=0;d<3;d++){
				v2 += vel[p+d]*(vel[p+d]+dv[d]);
				vel[p+d] += dv[d];
			}
			acc += v2;
		}
		if(kinEnergy){ kinEnergy[s] = acc; kinEnergy[s] *= 0.5*mass[s]; }
		orc_gmul(E,n,mass[s]/charge[s]);
	}
}

/* src/pusher.c:1233-1237 */
static void add_cross(const double *a, const double *b, double *res){
	res[0] +=  (a[1]*b[2]-a[2]*b[1]);
	res[1] += -(a[0]*b[2]-a[2]*b[0]);
	res[2] +=  (a[0]*b[1]-a[1]*b[0]);
}

/* src/pusher.c:394-483.  bugCompatible!=0 reproduces quirk Q3 (the rotation acts on
 * particle 0's velocity); bugCompatible==0 is the textbook Boris rotation of particle p,
 * which is what the product implements.  Both coincide when T=S=0 or for particle 0. */
void orc_boris3d1(double *pos, double *vel, int nSpecies, const long *iStart, const long *iStop,
                  const double *charge, const double *mass, double *E, const int *size,
                  const double *T, const double *S, double *kinEnergy, int bugCompatible){
	long sx = size[0], sy = size[1], n = 3L*size[0]*size[1]*size[2];
	for(int s=0;s<nSpecies;s++){
		orc_gmul(E,n,charge[s]/mass[s]);
		double acc = 0;
		for(long p=3*iStart[s]; p<3*iStop[s]; p+=3){
			double dv[3], vPrime[3];
			interp3d1(dv,&pos[p],E,sx,sy);
			for(int d=0;d<3;d++) vel[p+d] += 0.5*dv[d];
			double *v = bugCompatible ? vel : &vel[p];
			memcpy(vPrime,v,sizeof vPrime);
			add_cross(v,&T[3*s],vPrime);
			add_cross(vPrime,&S[3*s],v);
			double v2 = 0;
			for(int d=0;d<3;d++) v2 += pow(vel[p+d],2);
			acc += v2;
			for(int d=0;d<3;d++) vel[p+d] += 0.5*dv[d];
		}
		if(kinEnergy){ kinEnergy[s] = acc; kinEnergy[s] *= 0.5*mass[s]; }
		orc_gmul(E,n,mass[s]/charge[s]);
	}
}

/* src/pusher.c:485-505 */
void orc_rotation_parameters(int nSpecies, const double *BExt, const double *charge,
                             const double *mass, double *T, double *S){
	for(int s=0;s<nSpecies;s++){
		double factor = 0.5*charge[s]/mass[s];
		double denom = 1;
		for(int p=0;p<3;p++){ T[3*s+p] = factor*BExt[p]; denom += pow(T[3*s+p],2); }
		double mul = 2.0/denom;
		for(int p=0;p<3;p++) S[3*s+p] = mul*T[3*s+p];
	}
}

/* src/pusher.c:512-572.  gZero, then per species: rho *= 1/q, scatter raw weights, rho *= q (quirk Q1). */
void orc_distr3d1(const double *pos, int nSpecies, const long *iStart, const long *iStop,
                  const double *charge, double *rho, const int *size){
	long sx = size[0], sy = size[1], n = (long)size[0]*size[1]*size[2];
	for(long g=0;g<n;g++) rho[g] = 0;
	for(int s=0;s<nSpecies;s++){
		orc_gmul(rho,n,1.0/charge[s]);
		for(long i=iStart[s]; i<iStop[s]; i++){
			const double *r = &pos[3*i];
			int j = (int)r[0], k = (int)r[1], l = (int)r[2];
			double x = r[0]-j, y = r[1]-k, z = r[2]-l;
			double xc = 1-x, yc = 1-y, zc = 1-z;
			long p = IDX(j,k,l,sx,sy);
			rho[p]            += xc*yc*zc;
			rho[p+1]          += x *yc*zc;
			rho[p+sx]         += xc*y *zc;
			rho[p+sx+1]       += x *y *zc;
			rho[p+sx*sy]      += xc*yc*z;
			rho[p+sx*sy+1]    += x *yc*z;
			rho[p+sx*sy+sx]   += xc*y *z;
			rho[p+sx*sy+sx+1] += x *y *z;
		}
		orc_gmul(rho,n,charge[s]);
	}
}

/* The N-dimensional select() targets for nDims = 3 (src/pusher.c:215-391, 574-678, 1124-1180).  The reference recurses over the
 * dimensions (puInterpND1Inner :1147, puDistrND1Inner :626): the factor handed down is (z weight)*1, then (y weight)*that, and
 * the innermost level adds (x complement*factor)*value, then (x decimal*factor)*value, to the running result; corners are
 * visited z-stay before z-incr, y-stay before y-incr.  Written here as loops over (dz, dy) with the same operation order.
 * order 0 = nearest grid point, (int)(pos+0.5) (:1164, :644). */
void orc_acc_nd(double *pos, double *vel, int nSpecies, const long *iStart, const long *iStop,
                const double *charge, const double *mass, double *E, const int *size, int order, double *kinEnergy){
	long sx = size[0], sy = size[1], n = 3L*size[0]*size[1]*size[2];
	for(int s=0;s<nSpecies;s++){
		orc_gmul(E,n,charge[s]/mass[s]);
		double acc = 0;
		for(long p=3*iStart[s]; p<3*iStop[s]; p+=3){
			double dv[3] = {0,0,0};
			if(order==0){
				int j = (int)(pos[p]+0.5), k = (int)(pos[p+1]+0.5), l = (int)(pos[p+2]+0.5);
				for(int d=0;d<3;d++) dv[d] = E[3*IDX(j,k,l,sx,sy)+d];
			} else {
				int j = (int)pos[p], k = (int)pos[p+1], l = (int)pos[p+2];
				double dec[3] = {pos[p]-j, pos[p+1]-k, pos[p+2]-l};
				double com[3] = {1-dec[0], 1-dec[1], 1-dec[2]};
				for(int dz=0;dz<2;dz++){
					double fz = (dz ? dec[2] : com[2])*1.0;
					for(int dy=0;dy<2;dy++){
						double f = (dy ? dec[1] : com[1])*fz;
						long q = 3*IDX(j,k+dy,l+dz,sx,sy);
						for(int d=0;d<3;d++){
							dv[d] += com[0]*f*E[q+d];
							dv[d] += dec[0]*f*E[q+d+3];
						}
					}
				}
			}
			double v2 = 0;
			for(int d=0;d<3;d++){
				v2 += vel[p+d]*(vel[p+d]+dv[d]);
				vel[p+d] += dv[d];
			}
			acc += v2;
		}
		if(kinEnergy){ kinEnergy[s] = acc; kinEnergy[s] *= 0.5*mass[s]; }
		orc_gmul(E,n,mass[s]/charge[s]);
	}
}

void orc_distr_nd(const double *pos, int nSpecies, const long *iStart, const long *iStop,
                  const double *charge, double *rho, const int *size, int order){
	long sx = size[0], sy = size[1], n = (long)size[0]*size[1]*size[2];
	for(long g=0;g<n;g++) rho[g] = 0;
	for(int s=0;s<nSpecies;s++){
		orc_gmul(rho,n,1.0/charge[s]);
		for(long i=iStart[s]; i<iStop[s]; i++){
			const double *r = &pos[3*i];
			if(order==0){
				int j = (int)(r[0]+0.5), k = (int)(r[1]+0.5), l = (int)(r[2]+0.5);
				rho[IDX(j,k,l,sx,sy)]++;
				continue;
			}
			int j = (int)r[0], k = (int)r[1], l = (int)r[2];
			double dec[3] = {r[0]-j, r[1]-k, r[2]-l};
			double com[3] = {1-dec[0], 1-dec[1], 1-dec[2]};
			for(int dz=0;dz<2;dz++){
				double fz = (dz ? dec[2] : com[2])*1.0;
				for(int dy=0;dy<2;dy++){
					double f = (dy ? dec[1] : com[1])*fz;
					long q = IDX(j,k+dy,l+dz,sx,sy);
					rho[q]   += com[0]*f;
					rho[q+1] += dec[0]*f;
				}
			}
		}
		orc_gmul(rho,n,charge[s]);
	}
}

/* src/grid.c:1094-1099: upper thresholds are counted from the upper edge. */
void orc_thresholds(const int *size, const double *thrIn, double *thrOut){
	for(int d=0;d<3;d++){
		thrOut[d] = thrIn[d];
		thrOut[3+d] = (size[d]-1) - thrIn[3+d];
	}
}

/* src/pusher.c:782-855: serial classify, pack into the neighbour's buffer, back-fill the hole
 * with the species' last particle and re-examine the slot. */
void orc_extract3d(double *pos, double *vel, int nSpecies, const long *iStart, long *iStop,
                   const double *thr, double **emigrants, long *nEmigrants){
	double *cursor[27];
	for(int ne=0;ne<27;ne++) cursor[ne] = emigrants[ne];
	for(int i=0;i<27*nSpecies;i++) nEmigrants[i] = 0;
	for(int s=0;s<nSpecies;s++){
		long i = iStart[s];
		while(i<iStop[s]){
			double x = pos[3*i], y = pos[3*i+1], z = pos[3*i+2];
			int nx = -(x<thr[0]) + (x>=thr[3]);
			int ny = -(y<thr[1]) + (y>=thr[4]);
			int nz = -(z<thr[2]) + (z>=thr[5]);
			int ne = 13 + nx + 3*ny + 9*nz;
			if(ne==13){ i++; continue; }
			double *c = cursor[ne];
			c[0]=x; c[1]=y; c[2]=z; c[3]=vel[3*i]; c[4]=vel[3*i+1]; c[5]=vel[3*i+2];
			cursor[ne] += 6;
			nEmigrants[ne*nSpecies+s]++;
			long last = iStop[s]-1;
			for(int d=0;d<3;d++){ pos[3*i+d] = pos[3*last+d]; vel[3*i+d] = vel[3*last+d]; }
			iStop[s]--;
		}
	}
}

/* ------------------------------------------------------------------------------------------
 * Topology (src/grid.c:166-171, src/pusher.c:1181-1231)
 * ---------------------------------------------------------------------------------------- */
static void rank_to_sub(const OrcTopo *t, int rank, int *sub){
	for(int d=0;d<3;d++){ sub[d] = rank % t->nSub[d]; rank /= t->nSub[d]; }
}
static int sub_to_rank(const OrcTopo *t, const int *sub){
	return sub[0] + t->nSub[0]*(sub[1] + t->nSub[1]*sub[2]);
}
int orc_neighbor_to_rank(const OrcTopo *t, int rank, int ne){
	int sub[3], nb[3];
	rank_to_sub(t,rank,sub);
	for(int d=0;d<3;d++){
		int n = (ne%3)-1; ne /= 3;
		nb[d] = (sub[d]+n+t->nSub[d]) % t->nSub[d];
	}
	return sub_to_rank(t,nb);
}
int orc_neighbor_to_reciprocal(int ne){
	int rec = 0, pw = 1;
	for(int d=0;d<3;d++){ rec += (2-(ne%3))*pw; ne /= 3; pw *= 3; }
	return rec;
}
int orc_rank_to_neighbor(const OrcTopo *t, int rank, int other){
	int sub[3], ne = 0, pw = 1;
	rank_to_sub(t,rank,sub);
	for(int d=0;d<3;d++){
		int n = other % t->nSub[d];
		n = (n-sub[d]+1+t->nSub[d]) % t->nSub[d];
		other /= t->nSub[d];
		ne += n*pw; pw *= 3;
	}
	return ne;
}

/* src/pusher.c:914-1035.  Counts travel first (exchangeNMigrants), then the packed
 * (x,y,z,vx,vy,vz) records; the receiver shifts positions by (n_d)*trueSize[d] where n is
 * the direction the message came FROM (shiftImmigrants :941) and appends species by species
 * at iStop[s] (importParticles :967).  Import order: ascending receiver-side neighbour index. */
void orc_migrate(const OrcTopo *t, double **pos, double **vel, int nSpecies, long **iStop,
                 double ***emigrants, long **nEmigrants, long **nImmigrants){
	int R = t->nRanks;
	for(int r=0;r<R;r++)
		for(int ne=0;ne<27;ne++){
			if(ne==13){ for(int s=0;s<nSpecies;s++) nImmigrants[r][13*nSpecies+s] = 0; continue; }
			int src = orc_neighbor_to_rank(t,r,ne);
			int rec = orc_neighbor_to_reciprocal(ne);     /* r is neighbour `rec` of src */
			for(int s=0;s<nSpecies;s++) nImmigrants[r][ne*nSpecies+s] = nEmigrants[src][rec*nSpecies+s];
		}
	for(int r=0;r<R;r++)
		for(int ne=0;ne<27;ne++){
			if(ne==13) continue;
			int src = orc_neighbor_to_rank(t,r,ne);
			int rec = orc_neighbor_to_reciprocal(ne);
			const double *msg = emigrants[src][rec];
			double shift[3]; int q = ne;
			for(int d=0;d<3;d++){ int n = q%3-1; q /= 3; shift[d] = n*t->trueSize[d]; }
			for(int s=0;s<nSpecies;s++){
				long cnt = nImmigrants[r][ne*nSpecies+s];
				double *pp = &pos[r][3*iStop[r][s]], *vv = &vel[r][3*iStop[r][s]];
				for(long i=0;i<cnt;i++){
					for(int d=0;d<3;d++){ double v = msg[d]; v += shift[d]; pp[d] = v; }
					for(int d=0;d<3;d++) vv[d] = msg[3+d];
					msg += 6; pp += 3; vv += 3;
				}
				iStop[r][s] += cnt;
			}
		}
}

/* ------------------------------------------------------------------------------------------
 * Grid: halo exchange, reductions, finite differences
 * ---------------------------------------------------------------------------------------- */

/* A slice = every element whose coordinate along d equals `o` (all components, ghost rims
 * included), in memory order (src/grid.c:72-147). */
static void slice_get(double *buf, const double *val, const int *size, int nV, int d, int o){
	long ext[3] = {size[0],size[1],size[2]};
	long n = 0;
	for(long l=0;l<ext[2];l++){ if(d==2 && l!=o) continue;
	for(long k=0;k<ext[1];k++){ if(d==1 && k!=o) continue;
	for(long j=0;j<ext[0];j++){ if(d==0 && j!=o) continue;
		long g = nV*IDX(j,k,l,ext[0],ext[1]);
		for(int c=0;c<nV;c++) buf[n++] = val[g+c];
	}}}
}
static void slice_put(const double *buf, double *val, const int *size, int nV, int d, int o, int add){
	long ext[3] = {size[0],size[1],size[2]};
	long n = 0;
	for(long l=0;l<ext[2];l++){ if(d==2 && l!=o) continue;
	for(long k=0;k<ext[1];k++){ if(d==1 && k!=o) continue;
	for(long j=0;j<ext[0];j++){ if(d==0 && j!=o) continue;
		long g = nV*IDX(j,k,l,ext[0],ext[1]);
		for(int c=0;c<nV;c++){ if(add) val[g+c] += buf[n++]; else val[g+c] = buf[n++]; }
	}}}
}

/* src/grid.c:349-406: upper layer travels up and lands in the receiver's lower place, then
 * the lower layer travels down.  TOHALO: take size-2 / 1, place 0 / size-1.
 * FROMHALO: take size-1 / 0, place 1 / size-2. */
void orc_halo_dim(const OrcTopo *t, double **val, const int *size, int nV, int d, int op, int dir){
	int R = t->nRanks;
	long nSlice = (long)nV*size[0]*size[1]*size[2]/size[d];
	double *buf = malloc(sizeof(double)*nSlice*R);
	int upTake = size[d]-2+dir, upPlace = size[d]-1-dir, loTake = 1-dir, loPlace = dir;
	int upper[64], lower[64];
	for(int r=0;r<R;r++){
		int sub[3]; rank_to_sub(t,r,sub);
		int s0 = sub[d];
		sub[d] = (s0+1)%t->nSub[d];               upper[r] = sub_to_rank(t,sub);
		sub[d] = (s0-1+t->nSub[d])%t->nSub[d];    lower[r] = sub_to_rank(t,sub);
	}
	for(int r=0;r<R;r++) slice_get(buf+nSlice*r,val[r],size,nV,d,upTake);
	for(int r=0;r<R;r++) slice_put(buf+nSlice*lower[r],val[r],size,nV,d,loPlace,op);
	for(int r=0;r<R;r++) slice_get(buf+nSlice*r,val[r],size,nV,d,loTake);
	for(int r=0;r<R;r++) slice_put(buf+nSlice*upper[r],val[r],size,nV,d,upPlace,op);
	free(buf);
}
/* src/grid.c:340-347 */
void orc_halo(const OrcTopo *t, double **val, const int *size, int nV, int op, int dir){
	for(int d=0;d<3;d++) orc_halo_dim(t,val,size,nV,d,op,dir);
}

/* src/grid.c:804-847: true-grid sum, x innermost, one running sum per nesting level. */
double orc_sum_true(const double *val, const int *size){
	double s3 = 0;
	for(int l=1;l<size[2]-1;l++){
		double s2 = 0;
		for(int k=1;k<size[1]-1;k++){
			double s1 = 0;
			for(int j=1;j<size[0]-1;j++) s1 += val[IDX(j,k,l,size[0],size[1])];
			s2 += s1;
		}
		s3 += s2;
	}
	return s3;
}

/* src/grid.c:730-779: mean over the global true grid, subtracted from every element. */
void orc_neutralize(const OrcTopo *t, double **val, const int *size){
	double tot = 0;
	for(int r=0;r<t->nRanks;r++){ double mine = orc_sum_true(val[r],size); if(r==0) tot = mine; else tot += mine; }
	double avg = tot/((double)((size[0]-2)*(size[1]-2)*(size[2]-2))*t->nRanks);
	long n = (long)size[0]*size[1]*size[2];
	for(int r=0;r<t->nRanks;r++) for(long g=0;g<n;g++) val[r][g] -= avg;
}

/* src/grid.c:226-261 restricted to what survives the following halo set: true nodes only
 * (quirk Q9: the reference also writes junk into ghosts inside the flat range). */
void orc_findiff1st(const double *phi, double *E, const int *size){
	long sx = size[0], sy = size[1];
	long sp[3] = {1, sx, sx*sy};
	long start = sp[0]+sp[1]+sp[2], end = sx*sy*size[2]-start;
	for(int d=0;d<3;d++)
		for(long g=start; g<end; g++)
			E[3*g+d] = 0.5*(phi[g+sp[d]] - phi[g-sp[d]]);
}

/* src/grid.c:1276-1321 */
double orc_pot_energy(const double *rho, const double *phi, const int *size){
	double s3 = 0;
	for(int l=1;l<size[2]-1;l++){
		double s2 = 0;
		for(int k=1;k<size[1]-1;k++){
			double s1 = 0;
			for(int j=1;j<size[0]-1;j++){ long g = IDX(j,k,l,size[0],size[1]); s1 += rho[g]*phi[g]; }
			s2 += s1;
		}
		s3 += s2;
	}
	return 0.5*s3;
}

/* ------------------------------------------------------------------------------------------
 * Multigrid
 * ---------------------------------------------------------------------------------------- */

/* src/multigrid.c:683-767 (quirk Q8): per cycle, first the nodes with (j+k+l) odd, halo set,
 * neutralise (gBnd, periodic), then the even ones, halo set, neutralise.  The reference
 * also sweeps x/y ghost columns; those values are overwritten by the halo set and no true
 * node reads a same-colour node, so only true nodes are updated here. */
static void gs_colour(double *phi, const double *rho, const int *size, int parity){
	long sx = size[0], sy = size[1], sxy = sx*sy;
	const double coeff = 1./6.;
	for(int l=1;l<size[2]-1;l++) for(int k=1;k<size[1]-1;k++) for(int j=1;j<size[0]-1;j++){
		if(((j+k+l)&1)!=parity) continue;
		long g = IDX(j,k,l,sx,sy);
		phi[g] = coeff*( phi[g+1] + phi[g-1] + phi[g+sx] + phi[g-sx] + phi[g+sxy] + phi[g-sxy] + rho[g]);
	}
}
void orc_gs3d(const OrcTopo *t, double **phi, double **rho, const int *size, int nCycles){
	for(int c=0;c<nCycles;c++){
		for(int par=1;par>=0;par--){
			for(int r=0;r<t->nRanks;r++) gs_colour(phi[r],rho[r],size,par);
			orc_halo(t,phi,size,1,0,0);
			orc_neutralize(t,phi,size);
		}
	}
}

/* mgJacob3D (src/multigrid.c:500-551): all true nodes from the old values, then halo + gBnd (bnd == NULL: periodic) */
void orc_bnd(const OrcTopo *t, double **val, const int *size, const int *bnd, double **bndSlice);
void orc_jacobi3d(const OrcTopo *t, double **phi, double **rho, const int *size, int nCycles, const int *bnd, double **bndSlice){
	long sx = size[0], sxy = (long)size[0]*size[1], n = sxy*size[2];
	const double coeff = 1./6;
	double *tmp = malloc(sizeof(double)*n);
	for(int c=0;c<nCycles;c++){
		for(int r=0;r<t->nRanks;r++){
			for(int l=1;l<size[2]-1;l++) for(int k=1;k<size[1]-1;k++) for(int j=1;j<size[0]-1;j++){
				long g = IDX(j,k,l,size[0],size[1]);
				tmp[g] = coeff*(phi[r][g+1] + phi[r][g-1] + phi[r][g+sx] + phi[r][g-sx] + phi[r][g+sxy] + phi[r][g-sxy] + rho[r][g]);
			}
			for(int l=1;l<size[2]-1;l++) for(int k=1;k<size[1]-1;k++) for(int j=1;j<size[0]-1;j++){
				long g = IDX(j,k,l,size[0],size[1]);
				phi[r][g] = tmp[g];
			}
		}
		orc_halo(t,phi,size,1,0,0);
		if(bnd) orc_bnd(t,phi,size,bnd,bndSlice); else orc_neutralize(t,phi,size);
	}
	free(tmp);
}

/* ---- boundary conditions (src/grid.c:921-1023) -------------------------------------------------------------------
 * bnd[8]: bndType per boundary index as in Grid::bnd of a rank-4 grid (1..3 lower x,y,z; 5..7 upper; 0 and 4 unused):
 * 1 PERIODIC, 2 DIRICHLET, 3 NEUMANN.  bndSlice[r]: 8 slices of orc_slice_max() doubles, slice `boundary` in the element
 * order of getSlice/setSlice (src/grid.c:467, 940-952). */
long orc_slice_max(const int *size){
	/* nSliceMax of src/grid.c:455-465 runs over all rank dimensions INCLUDING the component axis, whose "slice" is the
	 * whole scalar grid: that term always wins */
	return (long)size[0]*size[1]*size[2];
}
/* src/grid.c:608-662 */
void orc_set_bnd_slices(const OrcTopo *t, int rank, const int *size, const int *bnd, double *bndSlice){
	long nMax = orc_slice_max(size);
	int sub[3]; rank_to_sub(t,rank,sub);
	for(int d=1;d<4;d++){
		if(sub[d-1]==0 && (bnd[d]==2 || bnd[d]==3)) for(long s=0;s<nMax;s++) bndSlice[s+nMax*d] = bnd[d]==2 ? 1. : 2.;
		if(sub[d-1]==t->nSub[d-1]-1 && (bnd[d+4]==2 || bnd[d+4]==3)) for(long s=0;s<nMax;s++) bndSlice[s+nMax*(d+4)] = bnd[d+4]==2 ? 1. : 2.;
	}
}
/* gDirichlet :929-956 (slice 1 on a lower edge, size-1 on an upper edge - the reference's own asymmetry) and
 * gNeumann :958-990 (ghost slice := slice two further in, minus twice the boundary value) */
static void edge(double *val, const int *size, int boundary, int kind, const double *bndSlice){
	int d = boundary%4 - 1, upper = boundary>4;
	long nMax = orc_slice_max(size), ns = (long)size[0]*size[1]*size[2]/size[d];
	const double *b = bndSlice + boundary*nMax;
	if(kind==2){ slice_put(b,val,size,1,d,1+upper*(size[d]-2),0); return; }
	int offset = upper*(size[d]-1);
	double *buf = malloc(sizeof(double)*ns);
	slice_get(buf,val,size,1,d,offset+2-4*upper);
	for(long s=0;s<ns;s++) buf[s] -= 2*b[s];
	slice_put(buf,val,size,1,d,offset,0);
	free(buf);
}
/* gBnd :992-1023 */
void orc_bnd(const OrcTopo *t, double **val, const int *size, const int *bnd, double **bndSlice){
	int periodic = 0;
	for(int d=1;d<4;d++) if(bnd[d]==1) periodic = 1;
	if(periodic) orc_neutralize(t,val,size);
	for(int r=0;r<t->nRanks;r++){
		int sub[3]; rank_to_sub(t,r,sub);
		for(int d=1;d<4;d++) if(sub[d-1]==0 && (bnd[d]==2 || bnd[d]==3)) edge(val[r],size,d,bnd[d],bndSlice[r]);
		for(int d=5;d<8;d++) if(sub[d-5]==t->nSub[d-5]-1 && (bnd[d]==2 || bnd[d]==3)) edge(val[r],size,d,bnd[d],bndSlice[r]);
	}
}
/* mgGS3D with gBnd after every colour (src/multigrid.c:683-767) */
void orc_gs3d_bnd(const OrcTopo *t, double **phi, double **rho, const int *size, int nCycles, const int *bnd, double **bndSlice){
	for(int c=0;c<nCycles;c++){
		for(int par=1;par>=0;par--){
			for(int r=0;r<t->nRanks;r++) gs_colour(phi[r],rho[r],size,par);
			orc_halo(t,phi,size,1,0,0);
			orc_bnd(t,phi,size,bnd,bndSlice);
		}
	}
}

/* src/multigrid.c:1385-1403 + src/grid.c:296-334, on true nodes (ghosts are set by the halo
 * exchange that always follows). */
void orc_residual(double *res, const double *rho, const double *phi, const int *size){
	long sx = size[0], sy = size[1], sxy = sx*sy;
	for(int l=1;l<size[2]-1;l++) for(int k=1;k<size[1]-1;k++) for(int j=1;j<size[0]-1;j++){
		long g = IDX(j,k,l,sx,sy);
		double r = -6.*phi[g];
		r += phi[g+1] + phi[g-1] + phi[g+sx] + phi[g-sx] + phi[g+sxy] + phi[g-sxy];
		r += rho[g];
		res[g] = r;
	}
}

/* src/multigrid.c:844-911: coarse (J,K,L) is centred on fine (2J-1,2K-1,2L-1). */
void orc_half_restrict3d(const double *f, const int *fs, double *c, const int *cs){
	long fx = fs[0], fxy = (long)fs[0]*fs[1];
	const double coeff = 1./12.;
	for(int L=1;L<cs[2]-1;L++) for(int K=1;K<cs[1]-1;K++) for(int J=1;J<cs[0]-1;J++){
		long g = IDX(2*J-1,2*K-1,2*L-1,fs[0],fs[1]);
		c[IDX(J,K,L,cs[0],cs[1])] = coeff*(6*f[g] + f[g+1] + f[g-1] + f[g+fx] + f[g-fx] + f[g+fxy] + f[g-fxy]);
	}
}

/* src/multigrid.c:1127-1238: inject at odd fine nodes, then fill even l, even k, even j by
 * midpoint averages, exchanging the halo of the dimension about to be interpolated. */
void orc_bilin_prol3d(const OrcTopo *t, double **fine, const int *fs, double **coarse, const int *cs){
	int R = t->nRanks;
	long fx = fs[0], fxy = (long)fs[0]*fs[1];
	for(int r=0;r<R;r++)
		for(int L=1;L<cs[2]-1;L++) for(int K=1;K<cs[1]-1;K++) for(int J=1;J<cs[0]-1;J++)
			fine[r][IDX(2*J-1,2*K-1,2*L-1,fs[0],fs[1])] = coarse[r][IDX(J,K,L,cs[0],cs[1])];
	orc_halo_dim(t,fine,fs,1,2,0,0);
	for(int r=0;r<R;r++)
		for(int l=2;l<fs[2]-1;l+=2) for(int k=1;k<fs[1]-1;k+=2) for(int j=1;j<fs[0]-1;j+=2){
			long g = IDX(j,k,l,fs[0],fs[1]);
			fine[r][g] = 0.5*(fine[r][g-fxy]+fine[r][g+fxy]);
		}
	orc_halo_dim(t,fine,fs,1,1,0,0);
	for(int r=0;r<R;r++)
		for(int l=1;l<fs[2]-1;l++) for(int k=2;k<fs[1]-1;k+=2) for(int j=1;j<fs[0]-1;j+=2){
			long g = IDX(j,k,l,fs[0],fs[1]);
			fine[r][g] = 0.5*(fine[r][g-fx]+fine[r][g+fx]);
		}
	orc_halo_dim(t,fine,fs,1,0,0,0);
	for(int r=0;r<R;r++)
		for(int l=1;l<fs[2]-1;l++) for(int k=1;k<fs[1]-1;k++) for(int j=2;j<fs[0]-1;j+=2){
			long g = IDX(j,k,l,fs[0],fs[1]);
			fine[r][g] = 0.5*(fine[r][g-1]+fine[r][g+1]);
		}
}

struct OrcMg {
	OrcTopo topo;
	int nLevels, nPre, nPost, nCoarse;
	int size[16][3];
	double **rho[16], **phi[16], **res[16];   /* [level][rank]; level 0 borrowed per call */
	int nonPeriodic, bnd[8];                  /* orc_mg_set_bnd: boundary types of every level */
	int smoother[3];                          /* orc_mg_set_smoothers: pre, post, coarse; 0 gaussSeidelRB, 1 jacobian */
	double **bndSlice[16];                    /* [level][rank], 8 slices each */
};

/* src/multigrid.c:128-206, 297-349: level q has trueSize/2^q, one ghost layer per side. */
OrcMg *orc_mg_alloc(const OrcTopo *t, int nLevels, int nPre, int nPost, int nCoarse){
	OrcMg *mg = calloc(1,sizeof *mg);
	mg->topo = *t; mg->nLevels = nLevels; mg->nPre = nPre; mg->nPost = nPost; mg->nCoarse = nCoarse;
	for(int q=0;q<nLevels;q++){
		for(int d=0;d<3;d++) mg->size[q][d] = t->trueSize[d]/(1<<q) + 2;
		long n = (long)mg->size[q][0]*mg->size[q][1]*mg->size[q][2];
		mg->rho[q] = calloc(t->nRanks,sizeof(double*));
		mg->phi[q] = calloc(t->nRanks,sizeof(double*));
		mg->res[q] = calloc(t->nRanks,sizeof(double*));
		if(q>0) for(int r=0;r<t->nRanks;r++){
			mg->rho[q][r] = calloc(n,sizeof(double));
			mg->phi[q][r] = calloc(n,sizeof(double));
			mg->res[q][r] = calloc(n,sizeof(double));
		}
	}
	return mg;
}
/* non-periodic edges: allocates (zeroed) boundary slices for every level of phi */
void orc_mg_set_bnd(OrcMg *mg, const int *bnd){
	mg->nonPeriodic = 0;
	for(int b=0;b<8;b++){ mg->bnd[b] = bnd[b]; if(b%4 && bnd[b]!=1) mg->nonPeriodic = 1; }
	for(int q=0;q<mg->nLevels;q++){
		if(mg->bndSlice[q]) continue;
		mg->bndSlice[q] = calloc(mg->topo.nRanks,sizeof(double*));
		for(int r=0;r<mg->topo.nRanks;r++) mg->bndSlice[q][r] = calloc(8*orc_slice_max(mg->size[q]),sizeof(double));
	}
}
double *orc_mg_bnd_slice(OrcMg *mg, int level, int rank){ return mg->bndSlice[level] ? mg->bndSlice[level][rank] : 0; }
/* src/multigrid.c:1314-1379 */
void orc_mg_restrict_bnd(OrcMg *mg){
	for(int q=0;q<mg->nLevels-1;q++){
		long nF = orc_slice_max(mg->size[q]), nC = orc_slice_max(mg->size[q+1]);
		for(int r=0;r<mg->topo.nRanks;r++)
			for(int d=1;d<8;d++){ if(d==4) continue; for(long s=0;s<nC;s++) mg->bndSlice[q+1][r][s+nC*d] = mg->bndSlice[q][r][2*s+nF*d]; }
	}
}
static void mg_bnd(OrcMg *mg, int level){
	if(mg->nonPeriodic) orc_bnd(&mg->topo,mg->phi[level],mg->size[level],mg->bnd,mg->bndSlice[level]);
	else orc_neutralize(&mg->topo,mg->phi[level],mg->size[level]);
}
void orc_mg_set_smoothers(OrcMg *mg, int pre, int post, int coarse){ mg->smoother[0] = pre; mg->smoother[1] = post; mg->smoother[2] = coarse; }
static void mg_gs(OrcMg *mg, int level, int nCycles, int which){
	if(mg->smoother[which] == 1){
		orc_jacobi3d(&mg->topo,mg->phi[level],mg->rho[level],mg->size[level],nCycles,mg->nonPeriodic ? mg->bnd : 0,mg->nonPeriodic ? mg->bndSlice[level] : 0);
		return;
	}
	if(mg->nonPeriodic) orc_gs3d_bnd(&mg->topo,mg->phi[level],mg->rho[level],mg->size[level],nCycles,mg->bnd,mg->bndSlice[level]);
	else orc_gs3d(&mg->topo,mg->phi[level],mg->rho[level],mg->size[level],nCycles);
}
void orc_mg_free(OrcMg *mg){
	for(int q=0;q<mg->nLevels;q++) if(mg->bndSlice[q]){ for(int r=0;r<mg->topo.nRanks;r++) free(mg->bndSlice[q][r]); free(mg->bndSlice[q]); }
	for(int q=0;q<mg->nLevels;q++){
		if(q>0) for(int r=0;r<mg->topo.nRanks;r++){ free(mg->rho[q][r]); free(mg->phi[q][r]); free(mg->res[q][r]); }
		free(mg->rho[q]); free(mg->phi[q]); free(mg->res[q]);
	}
	free(mg);
}
double *orc_mg_level(OrcMg *mg, int which, int level, int rank, int *sizeOut){
	for(int d=0;d<3;d++) sizeOut[d] = mg->size[level][d];
	return (which==0 ? mg->rho : which==1 ? mg->phi : mg->res)[level][rank];
}

/* src/multigrid.c:1496-1548 */
static void vcycle_top(OrcMg *mg, int level, int top);
static void vcycle(OrcMg *mg, int level){ vcycle_top(mg, level, 0); }
static void vcycle_top(OrcMg *mg, int level, int top){
	const OrcTopo *t = &mg->topo;
	int R = t->nRanks, bottom = mg->nLevels-1;
	const int *sz = mg->size[level];
	double **phi = mg->phi[level], **rho = mg->rho[level], **res = mg->res[level];
	long n = (long)sz[0]*sz[1]*sz[2];
	if(level==bottom){
		orc_halo(t,phi,sz,1,0,0);
		orc_halo(t,rho,sz,1,0,0);
		orc_neutralize(t,rho,sz);
		mg_gs(mg,level,mg->nCoarse,2);
		mg_bnd(mg,level);
		orc_bilin_prol3d(t,mg->res[level-1],mg->size[level-1],phi,sz);
		return;
	}
	orc_halo(t,rho,sz,1,0,0);
	orc_neutralize(t,rho,sz);
	mg_gs(mg,level,mg->nPre,0);
	for(int r=0;r<R;r++) orc_residual(res[r],rho[r],phi[r],sz);
	orc_halo(t,res,sz,1,0,0);
	for(int r=0;r<R;r++) orc_half_restrict3d(res[r],sz,mg->rho[level+1][r],mg->size[level+1]);
	vcycle_top(mg,level+1,top);
	for(int r=0;r<R;r++) for(long g=0;g<n;g++) phi[r][g] += res[r][g];
	orc_halo(t,phi,sz,1,0,0);
	mg_bnd(mg,level);
	mg_gs(mg,level,mg->nPost,1);
	mg_bnd(mg,level);
	if(level>top) orc_bilin_prol3d(t,mg->res[level-1],mg->size[level-1],phi,sz);
}
/* mgW (src/multigrid.c:1675-1683): two recursive V-cycles that meet at level bottom/2 */
void orc_mg_wcycle(OrcMg *mg, double **rho0, double **phi0, double **res0){
	for(int r=0;r<mg->topo.nRanks;r++){ mg->rho[0][r] = rho0[r]; mg->phi[0][r] = phi0[r]; mg->res[0][r] = res0[r]; }
	int bottom = mg->nLevels-1, middle = bottom/2;
	vcycle_top(mg,0,middle);
	vcycle_top(mg,middle,0);
}
/* mgVRegular (src/multigrid.c:1559-1650), call for call (it SUBTRACTS the prolonged correction) */
void orc_mg_vregular(OrcMg *mg, double **rho0, double **phi0, double **res0){
	const OrcTopo *t = &mg->topo;
	int R = t->nRanks, bottom = mg->nLevels-1;
	for(int r=0;r<R;r++){ mg->rho[0][r] = rho0[r]; mg->phi[0][r] = phi0[r]; mg->res[0][r] = res0[r]; }
	for(int cur=0;cur<bottom;cur++){
		const int *sz = mg->size[cur];
		long n = (long)sz[0]*sz[1]*sz[2];
		orc_halo(t,mg->phi[cur],sz,1,0,0);
		mg_bnd(mg,cur);
		orc_neutralize(t,mg->rho[cur],sz);
		mg_gs(mg,cur,mg->nPre,0);
		orc_halo(t,mg->rho[cur],sz,1,0,0);
		mg_bnd(mg,cur);
		for(int r=0;r<R;r++){ for(long g=0;g<n;g++) mg->res[cur][r][g] = 0; orc_residual(mg->res[cur][r],mg->rho[cur][r],mg->phi[cur][r],sz); }
		orc_halo(t,mg->res[cur],sz,1,0,0);
		for(int r=0;r<R;r++) orc_half_restrict3d(mg->res[cur][r],sz,mg->rho[cur+1][r],mg->size[cur+1]);
	}
	{
		const int *sz = mg->size[bottom];
		orc_neutralize(t,mg->rho[bottom],sz);
		orc_halo(t,mg->rho[bottom],sz,1,0,0);
		mg_gs(mg,bottom,mg->nCoarse,2);
		orc_halo(t,mg->phi[bottom],sz,1,0,0);
		mg_bnd(mg,bottom);
		orc_bilin_prol3d(t,mg->res[bottom-1],mg->size[bottom-1],mg->phi[bottom],sz);
	}
	for(int cur=bottom-1;cur>=0;cur--){
		const int *sz = mg->size[cur];
		long n = (long)sz[0]*sz[1]*sz[2];
		for(int r=0;r<R;r++) for(long g=0;g<n;g++) mg->phi[cur][r][g] -= mg->res[cur][r][g];
		orc_halo(t,mg->phi[cur],sz,1,0,0);
		mg_bnd(mg,cur);
		mg_gs(mg,cur,mg->nPost,1);
		mg_bnd(mg,cur);
		if(cur>0) orc_bilin_prol3d(t,mg->res[cur-1],mg->size[cur-1],mg->phi[cur],sz);
	}
}

void orc_mg_vcycle(OrcMg *mg, double **rho0, double **phi0, double **res0){
	for(int r=0;r<mg->topo.nRanks;r++){ mg->rho[0][r] = rho0[r]; mg->phi[0][r] = phi0[r]; mg->res[0][r] = res0[r]; }
	vcycle(mg,0);
}

/* src/multigrid.c:1688-1706 (nLevels>1 branch): V-cycles until the RMS residual over the
 * global true grid is <= tol; the residual grid is squared in place (mgSumTrueSquared :1471). */
int orc_mg_solve(OrcMg *mg, double **rho0, double **phi0, double **res0, double tol,
                 int maxCycles, double *barResOut, int cap){
	const OrcTopo *t = &mg->topo;
	const int *sz = mg->size[0];
	long n = (long)sz[0]*sz[1]*sz[2];
	double barRes = 2.;
	int cycles = 0;
	while(barRes>tol && cycles<maxCycles){
		orc_mg_vcycle(mg,rho0,phi0,res0);
		for(int r=0;r<t->nRanks;r++) orc_residual(res0[r],rho0[r],phi0[r],sz);
		orc_halo(t,res0,sz,1,0,0);
		double sum = 0;
		for(int r=0;r<t->nRanks;r++){
			for(long g=0;g<n;g++) res0[r][g] = res0[r][g]*res0[r][g];
			double mine = orc_sum_true(res0[r],sz);
			if(r==0) sum = mine; else sum += mine;
		}
		long tot = 1;
		for(int d=0;d<3;d++) tot *= (long)t->nSub[d]*(sz[d]-2);
		barRes = sum;
		barRes /= tot;
		barRes = sqrt(barRes);
		if(cycles<cap) barResOut[cycles] = barRes;
		cycles++;
	}
	return cycles;
}

/* ------------------------------------------------------------------------------------------
 * One time step in the canonical order (src/main.c:197-274 without object calls, one fold,
 * one solve, no HDF5: SURVEY 8c / quirks Q6, Q7).
 * ---------------------------------------------------------------------------------------- */
void orc_field_solve(OrcSim *s){
	const OrcTopo *t = &s->topo;
	int R = t->nRanks;
	long nE = 3L*s->size[0]*s->size[1]*s->size[2];
	for(int r=0;r<R;r++) orc_distr3d1(s->pos[r],s->nSpecies,s->iStart[r],s->iStop[r],s->charge,s->rho[r],s->size);
	orc_halo(t,s->rho,s->size,1,1,1);
	s->lastCycles = orc_mg_solve(s->mg,s->rho,s->phi,s->res,1e-10,1000,s->lastBarRes,256);
	orc_halo(t,s->phi,s->size,1,0,0);
	for(int r=0;r<R;r++) orc_findiff1st(s->phi[r],s->E[r],s->size);
	orc_halo(t,s->E,s->size,3,0,0);
	for(int r=0;r<R;r++) orc_gmul(s->E[r],nE,-1.);
}

void orc_accelerate(OrcSim *s, double scaleE){
	int R = s->topo.nRanks;
	long nE = 3L*s->size[0]*s->size[1]*s->size[2];
	double ke[8];
	for(int q=0;q<=s->nSpecies;q++) s->kinEnergy[q] = 0;
	for(int r=0;r<R;r++){
		if(scaleE!=1.0) orc_gmul(s->E[r],nE,scaleE);
		orc_acc3d1(s->pos[r],s->vel[r],s->nSpecies,s->iStart[r],s->iStop[r],s->charge,s->mass,s->E[r],s->size,ke);
		if(scaleE!=1.0) orc_gmul(s->E[r],nE,1.0/scaleE);
		for(int q=0;q<s->nSpecies;q++) s->kinEnergy[q] += ke[q];
	}
	for(int q=0;q<s->nSpecies;q++) s->kinEnergy[s->nSpecies] += s->kinEnergy[q];
}

void orc_step(OrcSim *s){
	const OrcTopo *t = &s->topo;
	int R = t->nRanks;
	for(int r=0;r<R;r++) orc_move(s->pos[r],s->vel[r],s->nSpecies,s->iStart[r],s->iStop[r]);
	for(int r=0;r<R;r++) orc_extract3d(s->pos[r],s->vel[r],s->nSpecies,s->iStart[r],s->iStop[r],s->thresholds,s->emigrants[r],s->nEmigrants[r]);
	orc_migrate(t,s->pos,s->vel,s->nSpecies,s->iStop,s->emigrants,s->nEmigrants,s->nImmigrants);
	orc_field_solve(s);
	orc_accelerate(s,1.0);
	double pe = 0;
	for(int r=0;r<R;r++) pe += orc_pot_energy(s->rho[r],s->phi[r],s->size);
	s->potEnergy = pe;
}

/* ------------------------------------------------------------------------------------------
 * pNew / pCut (src/population.c:430-466): the reference's single-particle insert and the
 * extract-with-back-fill that puExtractEmigrants3D's serial loop is built on; pinned by
 * test/population.test.c:10-56.  p is the FLAT index of the particle's first coordinate.
 * ---------------------------------------------------------------------------------------- */
int orc_pnew(double *pos, double *vel, const long *iStart, long *iStop, int s, const double *p3, const double *v3){
	if(iStop[s] >= iStart[s+1]) return 0;                 /* "New particle ignored" */
	long p = iStop[s]*3;
	for(int d=0;d<3;d++){ pos[p+d] = p3[d]; vel[p+d] = v3[d]; }
	iStop[s]++;
	return 1;
}
void orc_pcut(double *pos, double *vel, long *iStop, int s, long p, double *p3, double *v3){
	long pLast = (iStop[s]-1)*3;
	for(int d=0;d<3;d++){
		p3[d] = pos[p+d]; v3[d] = vel[p+d];
		pos[p+d] = pos[pLast+d]; vel[p+d] = vel[pLast+d];
	}
	iStop[s]--;
}
