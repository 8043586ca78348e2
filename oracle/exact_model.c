// TEST INFRASTRUCTURE ONLY (design model, not shipped, not called by the product).
// CPU model of the order-exact PARALLEL formulation of UpdateESDF used by fiesta_b200/csrc/fb_exact.cu (multi-version
// fixpoint per FIFO generation, timestamps = (queue position, direction)), checked against the sequential oracle in the
// same process: `gcc -O2 -ffp-contract=off -o /tmp/exact_model oracle/exact_model.c -lm && /tmp/exact_model 32 0.7 6 1500 1`
// prints, per update, the reference's expansion count, ours, and the number of distance / closest-obstacle mismatches
// (all 0).  Every "for each element / voxel" loop below is data-parallel.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "esdf_oracle.c"
typedef unsigned int u32; typedef unsigned long long u64;
#define NONE 0xffffffffu
#define CU 0u   /* unknown */
#define CI 1u   /* inf */
static int GX,GY,GZ; static long N;
static u32 *C;            // codes
static u32 *M;            // element index of the live entry in the current generation
static u64 *LS;           // link sequence (time of last relink)
static u64 tclock=1;
static omap* O;
static inline u32 pack(int x,int y,int z){return ((u32)(x+1)<<20)|((u32)y<<10)|(u32)z;}
static inline void unpack(u32 c,int*x,int*y,int*z){*x=(int)(c>>20)-1;*y=(c>>10)&1023;*z=c&1023;}
static inline long vi(int x,int y,int z){return ((long)x*GY+y)*GZ+z;}
static inline int inb(int x,int y,int z){return x>=O->min_vec[0]&&x<=O->max_vec[0]&&y>=O->min_vec[1]&&y<=O->max_vec[1]&&z>=O->min_vec[2]&&z<=O->max_vec[2];}
static inline u32 d2(int x,int y,int z,u32 c){int ox,oy,oz;unpack(c,&ox,&oy,&oz);ox-=x;oy-=y;oz-=z;return (u32)(ox*ox+oy*oy+oz*oz);}
static inline int existc(u32 c){int x,y,z;unpack(c,&x,&y,&z);return O->occ[vi(x,y,z)]>O->l_occ;}
#define DINF 0xffffffffu
static inline u32 dist_of(long v,u32 c){ if(c<2) return DINF; int x=v/(GY*GZ),y=(v/GZ)%GY,z=v%GZ; return d2(x,y,z,c);}

typedef struct {u32 kind; u32 code;} beh_t; // kind 0 dead,1 pull,2 push
static u32 *E; static long nE; static beh_t *B0,*B1;

typedef struct {u32 d,c,ts;} st_t;
// State of voxel v as seen at time T (exclusive), given behaviours B
static st_t state_at(long v,u32 T,const beh_t*B){
  st_t s; s.c=C[v]; s.d=dist_of(v,s.c); s.ts=NONE; u32 d0=s.d;
  if(s.c==CU) return s;  // unknown voxels never accept
  int x=v/(GY*GZ),y=(v/GZ)%GY,z=v%GZ;
  if(!inb(x,y,z)) return s;   // pushes only go to in-box voxels
  for(int k=0;k<24;k++){ int qx=x-DIRS[k][0],qy=y-DIRS[k][1],qz=z-DIRS[k][2];
    if(qx<0||qy<0||qz<0||qx>=GX||qy>=GY||qz>=GZ) continue;
    u32 j=M[vi(qx,qy,qz)]; if(j==NONE) continue; if(B[j].kind!=2) continue; u32 ts=j*32+k; if(ts>=T) continue;
    u32 c=B[j].code; u32 d=d2(x,y,z,c); if(d<d0 && (d<s.d || (d==s.d && ts<s.ts))){s.d=d;s.c=c;s.ts=ts;} }
  u32 j=M[v]; if(j!=NONE && B[j].kind==1){ u32 ts=j*32+24; if(ts<T){ u32 c=B[j].code; u32 d=d2(x,y,z,c); if(d<d0 && (d<s.d||(d==s.d&&ts<s.ts))){s.d=d;s.c=c;s.ts=ts;} } }
  return s;
}
static long expansions; static long totrounds=0, totgens=0, maxrounds=0;
static void relax(void){
  while(nE){
    // initial guess: every element pushes its snapshot code
    for(long i=0;i<nE;i++){ B0[i].kind=2; B0[i].code=C[E[i]]; }
    int rounds=0;
    for(;;){ long changed=0; rounds++;
      for(long i=0;i<nE;i++){ long p=E[i]; int x=p/(GY*GZ),y=(p/GZ)%GY,z=p%GZ; u32 T0=i*32;
        st_t s=state_at(p,T0,B0); beh_t nb;
        u32 d0=dist_of(p,C[p]);
        if(s.d!=d0){ nb.kind=0; nb.code=0; }
        else { u32 curd=s.d, curc=s.c; int ch=0;
          for(int k=0;k<24;k++){ int nx=x+DIRS[k][0],ny=y+DIRS[k][1],nz=z+DIRS[k][2]; if(!inb(nx,ny,nz)) continue;
            st_t sn=state_at(vi(nx,ny,nz),T0,B0); if(sn.c<2) continue; u32 t=d2(x,y,z,sn.c); if(curd>t){curd=t;curc=sn.c;ch=1;} }
          if(ch){nb.kind=1;nb.code=curc;} else {nb.kind=2;nb.code=s.c;} }
        if(nb.kind!=B0[i].kind||nb.code!=B0[i].code) changed++;
        B1[i]=nb; }
      beh_t*t=B0;B0=B1;B1=t;
      totrounds+=rounds>0?0:0; if(!changed) break; if(rounds>10000){printf("no convergence\n");exit(1);} }
    totgens++; totrounds+=rounds; if(rounds>maxrounds)maxrounds=rounds;
    // commit: winners -> next generation in ts order
    long cap=nE*32; u32 *slotv=malloc(sizeof(u32)*cap); u32*slotc=malloc(sizeof(u32)*cap); for(long s=0;s<cap;s++) slotv[s]=NONE;
    for(long i=0;i<nE;i++){ if(B0[i].kind) expansions++; long p=E[i]; int x=p/(GY*GZ),y=(p/GZ)%GY,z=p%GZ;
      if(B0[i].kind==2){ u32 c=B0[i].code; for(int k=0;k<24;k++){ int nx=x+DIRS[k][0],ny=y+DIRS[k][1],nz=z+DIRS[k][2]; if(nx<0||ny<0||nz<0||nx>=GX||ny>=GY||nz>=GZ) continue; if(!inb(nx,ny,nz)) continue; long n=vi(nx,ny,nz);
          st_t f=state_at(n,NONE,B0); if(f.ts==(u32)(i*32+k)){ slotv[i*32+k]=(u32)n; slotc[i*32+k]=f.c; } } }
      else if(B0[i].kind==1){ st_t f=state_at(p,NONE,B0); if(f.ts==(u32)(i*32+24)){ slotv[i*32+24]=(u32)p; slotc[i*32+24]=f.c; } } }
    for(long i=0;i<nE;i++) M[E[i]]=NONE;
    long n2=0; u32*E2=malloc(sizeof(u32)*(cap?cap:1));
    for(long s=0;s<cap;s++) if(slotv[s]!=NONE){ u32 v=slotv[s]; C[v]=slotc[s]; LS[v]=tclock+s; M[v]=n2; E2[n2++]=v; }
    tclock+=cap+1;
    free(slotv);free(slotc); memcpy(E,E2,sizeof(u32)*n2); free(E2); nE=n2;
  }
}
typedef struct {u64 k1,k2; u32 v;} dep_t;
static int cmpdep(const void*a,const void*b){const dep_t*x=a,*y=b; if(x->k1!=y->k1) return x->k1<y->k1?-1:1; if(x->k2!=y->k2) return x->k2>y->k2?-1:1; return 0;}
int main(int argc,char**argv){
  int G=argc>1?atoi(argv[1]):24; double obs=argc>2?atof(argv[2]):0.7; int rounds=argc>3?atoi(argv[3]):6; int nops=argc>4?atoi(argv[4]):600; srand(argc>5?atoi(argv[5]):1);
  double org[3]={0,0,0},sz[3]={G*0.1-0.05,G*0.1-0.05,G*0.1-0.05}; O=fiesta_oracle_create(org,0.1,sz); fiesta_oracle_set_parameters(O,0.97,0.03,0.30,0.90,0.80);
  GX=O->gs[0];GY=O->gs[1];GZ=O->gs[2];N=(long)GX*GY*GZ; C=calloc(N,4); M=malloc(N*4); LS=calloc(N,8); for(long i=0;i<N;i++)M[i]=NONE;
  E=malloc(sizeof(u32)*N*32); B0=malloc(sizeof(beh_t)*N*4); B1=malloc(sizeof(beh_t)*N*4);
  for(int r=0;r<rounds;r++){
    int nev=r==0?(int)(N*obs):nops;
    for(int e=0;e<nev;e++){int v[3]={rand()%GX,rand()%GY,rand()%GZ}; fiesta_oracle_set_occupancy_vox(O,v,r==0?(rand()%50==0):rand()%2);}
    double*pre=malloc(N*8); memcpy(pre,O->occ,N*8); double*pred=malloc(N*8); memcpy(pred,O->dist,N*8);
    // queue orders come from the oracle's own queues (integration order is validated separately)
    fiesta_oracle_update_occupancy(O,1);
    long nins=fifo_size(&O->q_ins), ndel=fifo_size(&O->q_del);
    u32*ins=malloc(4*(nins+1)),*del=malloc(4*(ndel+1));
    for(long i=0;i<nins;i++){int*v=O->q_ins.e[O->q_ins.head+i].v; ins[i]=vi(v[0],v[1],v[2]);}
    for(long i=0;i<ndel;i++){int*v=O->q_del.e[O->q_del.head+i].v; del[i]=vi(v[0],v[1],v[2]);}
    for(long i=0;i<N;i++) if(pred[i]<0&&O->dist[i]>=0&&C[i]==CU) C[i]=CI;
    expansions=0;
    // E1 insert seeds, in order
    nE=0; for(long i=0;i<nins;i++){ long x=ins[i]; if(O->occ[x]>O->l_occ){ int a=x/(GY*GZ),b=(x/GZ)%GY,c=x%GZ; C[x]=pack(a,b,c); LS[x]=tclock++; M[x]=nE; E[nE++]=x; } }
    // E2 delete: dependants by dense scan, ordered by (delete rank, descending link time)
    u32*rank=malloc(4*N); for(long i=0;i<N;i++)rank[i]=NONE; long nd=0; for(long i=0;i<ndel;i++){ long x=del[i]; if(!(O->occ[x]>O->l_occ) && rank[x]==NONE) rank[x]=nd++; }
    long ndep=0; dep_t*deps=malloc(sizeof(dep_t)*N);
    for(long u=0;u<N;u++){ if(C[u]>=2){ int ox,oy,oz; unpack(C[u],&ox,&oy,&oz); long xo=vi(ox,oy,oz); if(rank[xo]!=NONE){ deps[ndep].k1=rank[xo]; deps[ndep].k2=LS[u]; deps[ndep].v=u; ndep++; } } }
    qsort(deps,ndep,sizeof(dep_t),cmpdep);
    u32*ord=malloc(4*N); for(long i=0;i<N;i++)ord[i]=NONE; for(long i=0;i<ndep;i++)ord[deps[i].v]=i;
    u32*nc0=malloc(4*(ndep+1)),*nc1=malloc(4*(ndep+1)); for(long i=0;i<ndep;i++)nc0[i]=CI;
    for(int it=0;;it++){ long ch=0;
      for(long i=0;i<ndep;i++){ long u=deps[i].v; int x=u/(GY*GZ),y=(u/GZ)%GY,z=u%GZ; u32 res=CI;
        for(int k=0;k<24;k++){ int nx=x+DIRS[k][0],ny=y+DIRS[k][1],nz=z+DIRS[k][2]; if(!inb(nx,ny,nz)) continue; long n=vi(nx,ny,nz); u32 c;
          if(ord[n]!=NONE){ if(ord[n]<(u32)i) c=nc0[ord[n]]; else continue; } else c=C[n];
          if(c>=2 && existc(c)){ res=c; break; } }
        nc1[i]=res; if(res!=nc0[i]) ch++; }
      u32*t=nc0;nc0=nc1;nc1=t; if(!ch) break; }
    for(long i=0;i<ndep;i++){ long u=deps[i].v; C[u]=nc0[i]; LS[u]=tclock++; if(nc0[i]>=2){ M[u]=nE; E[nE++]=u; } }
    fiesta_oracle_update_esdf(O);
    relax();
    long dm=0,cm=0; for(long i=0;i<N;i++){ u32 c=C[i]; double d; if(c==CU)d=-10000; else if(c==CI)d=10000; else {int x=i/(GY*GZ),y=(i/GZ)%GY,z=i%GZ; d=sqrt((double)d2(x,y,z,c))*0.1;}
      if(d!=O->dist[i])dm++; int ox=-10000,oy=-10000,oz=-10000; if(c>=2)unpack(c,&ox,&oy,&oz); if(ox!=O->cobs[3*i]||oy!=O->cobs[3*i+1]||oz!=O->cobs[3*i+2])cm++; }
    printf("[gens %ld rounds %ld max %ld] round %d ins %ld del %ld dep %ld | ref expansions %ld ours %ld | dist mismatches %ld cobs mismatches %ld\n",totgens,totrounds,maxrounds,r,nins,ndel,ndep,O->st_exp,expansions,dm,cm);
    free(pre);free(pred);free(ins);free(del);free(rank);free(deps);free(ord);free(nc0);free(nc1);
  }
  return 0;
}
