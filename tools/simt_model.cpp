// simt_model.cpp — SIMT cost model of wf_extend's node walk (development aid).
//
// Walks the REAL mode-2 layouts of Book-1 with REAL rays (tools/simt_model_dump.py writes both to /tmp/sim/: camera rays
// in 8x4-pixel warps, bounce rays from the oracle's scatter, shuffled), 32 lanes at a time, and charges warp-instructions
// per control-flow scheme:  0 = if/else loop, 1 = while-while with K parked leaves (K = 1: what the compiler makes of the
// plain loop), 2 / 3 = persistent warps with lane refill + parked leaves (exit rule: parked-lane / walker thresholds).
//   g++ -O2 -o sim tools/simt_model.cpp && ./sim /tmp/sim/b1.bin <scheme> [K] [instr per node visit] [threshold]
// It predicted the lane counts of wf_extend_stream well (26 modelled, 20.5 measured) and its overheads badly.
// SIMT cost model of wf_extend's walk: warps of 32 rays (same octant), real mode-2 layout, f32 exact slab test.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>
#include <algorithm>
#include <cstdint>
struct Node { float f[8]; };
static std::vector<Node> L[8];
static uint32_t NN;
struct Ray { float o[3], d[3], time, pad; };
static inline uint32_t fb(float x){ uint32_t u; memcpy(&u,&x,4); return u; }
struct Lane {
    Ray r; float inv[3]; float a; uint32_t i; bool done; float best; int obj;
    uint32_t pend[8]; int np;
    long visits, leaves;
};
static bool slab_miss(const Node& n, const Lane& l, float tmax) {
    float lo = 0.001f, hi = tmax;
    for (int k = 0; k < 3; ++k) {
        float t0 = (n.f[k] - l.r.o[k]) * l.inv[k], t1 = (n.f[4+k] - l.r.o[k]) * l.inv[k];
        if (t0 == t0) lo = std::fmax(lo, t0);
        if (t1 == t1) hi = std::fmin(hi, t1);
    }
    return hi <= lo;
}
// returns cost class: 0 = disc<0 early out, 1 = full path
static int sphere(Lane& l, const Node& n) {
    uint32_t meta = fb(n.f[3]); uint32_t kind = meta >> 30;
    float c[3] = {n.f[0], n.f[1], n.f[2]};
    if (kind == 2) for (int k=0;k<3;++k) c[k] += l.r.time * n.f[4+k];
    float oc[3]; for (int k=0;k<3;++k) oc[k] = l.r.o[k]-c[k];
    float hb = oc[0]*l.r.d[0]+oc[1]*l.r.d[1]+oc[2]*l.r.d[2];
    float cc = oc[0]*oc[0]+oc[1]*oc[1]+oc[2]*oc[2] - n.f[7]*n.f[7];
    float disc = hb*hb - l.a*cc;
    if (disc < 0) return 0;
    float s = std::sqrt(disc);
    float root = (-hb - s)/l.a;
    if (!(0.001f < root && root < l.best)) { root = (-hb + s)/l.a; if (!(0.001f < root && root < l.best)) return 1; }
    l.best = root; l.obj = meta & 0x3fffffff; return 1;
}
int main(int argc, char** argv) {
    const char* rayfile = argv[1]; int scheme = atoi(argv[2]); int K = argc > 3 ? atoi(argv[3]) : 1;
    int C_NODE = argc > 4 ? atoi(argv[4]) : 12; int C_LEAF0 = 30, C_LEAF1 = 62;
    FILE* f = fopen("/tmp/sim/layout.bin","rb"); fseek(f,0,SEEK_END); long sz=ftell(f); fseek(f,0,SEEK_SET);
    NN = sz/8/32; for (int o=0;o<8;++o){ L[o].resize(NN); fread(L[o].data(),32,NN,f);} fclose(f);
    f = fopen(rayfile,"rb"); fseek(f,0,SEEK_END); sz=ftell(f); fseek(f,0,SEEK_SET);
    std::vector<Ray> rays(sz/32); fread(rays.data(),32,rays.size(),f); fclose(f);
    // bin by octant (keeping order), warps of 32
    std::vector<Ray> q[8];
    for (auto& r: rays){ int oc=0; for(int k=0;k<3;++k){ uint32_t b=fb(r.d[k])-0x80000000u; if (b<0x7f800000u) oc|=1<<k;} q[oc].push_back(r);}    
    double cost=0, lane_iters=0, warp_iters=0, nrays=0, visits=0, leaves=0, outer=0, leafphase=0, leaflanes=0;
    int T = argc > 5 ? atoi(argv[5]) : 24;
    if (scheme == 2 || scheme == 3) for (int oc=0;oc<8;++oc) {
        // persistent warps: each warp streams through a contiguous slice of SLICE rays of the octant queue
        const size_t SLICE = 640; const std::vector<Node>& N = L[oc];
        for (size_t w0=0; w0+SLICE<=q[oc].size(); w0+=SLICE) {
            Lane ln[32]; size_t next=w0, end=w0+SLICE; bool live[32];
            auto load=[&](Lane& l){ l.r=q[oc][next++]; for(int k=0;k<3;++k) l.inv[k]=1.0f/l.r.d[k];
                l.a=l.r.d[0]*l.r.d[0]+l.r.d[1]*l.r.d[1]+l.r.d[2]*l.r.d[2]; l.i=0; l.done=false; l.best=INFINITY; l.obj=-1; l.np=0; };
            for(int t=0;t<32;++t){ load(ln[t]); live[t]=true; ln[t].visits=ln[t].leaves=0; }
            nrays+=SLICE;
            for(;;){
                // inner: walk until >= T lanes have a pending leaf, or any lane is full, or nobody can walk
                for(;;){ int act=0, havep=0; bool full=false;
                    for(int t=0;t<32;++t){ Lane& l=ln[t]; if(!live[t]) continue; if (l.np>0) ++havep; if (l.np>=K) full=true; }
                    int walkers=0; for(int t=0;t<32;++t) if(live[t] && !ln[t].done && ln[t].np<K) ++walkers;
                    if (scheme==3) { bool rem = next<end; if (walkers < (rem? T:1) ) break; } else
                    if (havep>=T || full || walkers==0) break;
                    // lanes that are done with pending 0 are idle until the refill point: also break if many idle
                    int idle=0; for(int t=0;t<32;++t) if(live[t] && ln[t].done && ln[t].np==0) ++idle;
                    if (scheme==2 && idle >= 32-T+1 && next<end) break;
                    for (int t=0;t<32;++t){ Lane& l=ln[t]; if (!live[t]||l.done||l.np>=K) continue; ++act; const Node& n=N[l.i]; uint32_t m=fb(n.f[3]);
                        if (m < (1u<<30)) { ++l.visits; l.i = slab_miss(n,l,l.best)? m : l.i+1; }
                        else if (m==0xffffffffu) { l.done=true; }
                        else { l.pend[l.np++]=l.i; l.i++; } }
                    warp_iters++; lane_iters+=act; cost += C_NODE + 2; }
                // one leaf phase: oldest pending of every lane
                { int lc=-1,cnt=0; for(int t=0;t<32;++t){ Lane& l=ln[t]; if(!live[t]||l.np==0) continue; ++cnt; ++l.leaves; lc=std::max(lc,sphere(l,N[l.pend[0]])); for(int j=1;j<l.np;++j) l.pend[j-1]=l.pend[j]; --l.np; }
                  if (lc>=0){ cost += (lc?C_LEAF1:C_LEAF0) + 6; leafphase++; leaflanes+=cnt; } }
                outer++;
                // retire + refill
                bool any=false, refilled=false;
                for(int t=0;t<32;++t){ Lane& l=ln[t]; if(!live[t]) continue; if (l.done && l.np==0){ visits+=l.visits; leaves+=l.leaves; l.visits=l.leaves=0; if (next<end){ load(l); refilled=true; } else live[t]=false; } if(live[t]) any=true; }
                if (refilled) cost += 14;
                cost += 4;
                if(!any) break;
            }
        }
    }
    if (scheme == 5) {
        // trip-count utilisation (mean / max of the per-ray visit counts of a warp) if a visit covered k binary nodes:
        // a wider node shortens every walk by the same factor and leaves the ratio where it is
        double sum[3] = {0,0,0}, mx[3] = {0,0,0}; const int ks[3] = {1, 2, 4};
        for (int oc=0;oc<8;++oc) for (size_t w=0; w+32<=q[oc].size(); w+=32) {
            long m[3] = {0,0,0};
            for (int t=0;t<32;++t){ Lane l; l.r=q[oc][w+t]; for(int k=0;k<3;++k) l.inv[k]=1.0f/l.r.d[k];
                l.a=l.r.d[0]*l.r.d[0]+l.r.d[1]*l.r.d[1]+l.r.d[2]*l.r.d[2]; l.i=0; l.best=INFINITY; l.obj=-1; long v=0;
                for(;;){ const Node& n=L[oc][l.i]; uint32_t mm=fb(n.f[3]); if (mm < (1u<<30)) { ++v; l.i = slab_miss(n,l,l.best)? mm : l.i+1; } else if (mm==0xffffffffu) break; else { sphere(l,n); ++v; l.i++; } }
                for (int j=0;j<3;++j){ long u=(v+ks[j]-1)/ks[j]; sum[j]+=u; m[j]=std::max(m[j],u); } }
            for (int j=0;j<3;++j) mx[j]+=m[j]; }
        for (int j=0;j<3;++j) printf("%s k=%d: mean walk %.1f visits, longest in its warp %.1f, utilisation %.3f\n", rayfile, ks[j], sum[j]/ (mx[j]? 1:1) / (double)(rays.size()/32*32), mx[j]/(rays.size()/32), sum[j]/(32.0*mx[j]));
        return 0;
    }
    if (scheme == 6) {
        // PAIR visits: an interior node's two child boxes are tested in one visit (two independent
        // chains behind one fetch), near child first (the per-octant layouts are already in near-first order), the far
        // child on a per-lane stack; leaves are tested in phases like scheme 1.  Children of the box at i: A = i + 1 and
        // B = skip link of A (pre-order); a box whose next entry is a leaf is that leaf's own box.
        const int C_PAIR = argc > 4 ? atoi(argv[4]) : 25;
        double pair_iters = 0, pair_lane = 0, boxtests = 0;
        for (int oc=0;oc<8;++oc) for (size_t w=0; w+32<=q[oc].size(); w+=32) {
            const std::vector<Node>& N = L[oc];
            struct PL { Lane l; std::vector<uint32_t> st; int64_t pend; bool done; };
            std::vector<PL> ln(32);
            // the first entries are the huge objects' leaves (no box) and then the root's first subtree
            uint32_t first = 0; while (fb(N[first].f[3]) >= (1u<<30) && fb(N[first].f[3]) != 0xffffffffu) ++first;
            for (int t=0;t<32;++t){ Lane& l=ln[t].l; l.r=q[oc][w+t]; for(int k=0;k<3;++k) l.inv[k]=1.0f/l.r.d[k];
                l.a=l.r.d[0]*l.r.d[0]+l.r.d[1]*l.r.d[1]+l.r.d[2]*l.r.d[2]; l.best=INFINITY; l.obj=-1; l.visits=l.leaves=0;
                for (uint32_t h=0; h<first; ++h) { sphere(l, N[h]); }           // huge leaves: converged, cheap
                ln[t].st.clear(); ln[t].pend=-1; ln[t].done=false;
                // the two root subtrees: treated as the children of a virtual root
                const uint32_t A=first, B=fb(N[first].f[3]);
                bool hb = fb(N[B].f[3]) != 0xffffffffu && !slab_miss(N[B], l, l.best); bool ha = !slab_miss(N[A], l, l.best); boxtests+=2;
                if (hb) ln[t].st.push_back(B); if (ha) ln[t].st.push_back(A);
            }
            nrays+=32; cost += 12;  // the huge leaves + virtual root, all lanes converged
            for(;;){
                bool any=false; for(int t=0;t<32;++t) if(!ln[t].done) any=true; if(!any) break;
                outer++;
                for(;;){ int act=0;
                    for (int t=0;t<32;++t){ PL& p=ln[t]; if (p.done||p.pend>=0) continue;
                        if (p.st.empty()) { p.done=true; continue; }
                        ++act; const uint32_t i=p.st.back(); p.st.pop_back();   // a box already known to be hit when pushed
                        // re-test against the current best (the pushed box may have been overtaken by a nearer hit)
                        if (slab_miss(N[i], p.l, p.l.best)) { ++boxtests; continue; }
                        const uint32_t m1=fb(N[i+1].f[3]);
                        if (m1 >= (1u<<30)) { p.pend = i+1; continue; }        // i is a leaf's own box
                        const uint32_t A=i+1, B=fb(N[A].f[3]);
                        const bool ha=!slab_miss(N[A], p.l, p.l.best), hb=!slab_miss(N[B], p.l, p.l.best); boxtests+=2; ++p.l.visits;
                        if (hb) p.st.push_back(B); if (ha) p.st.push_back(A);
                    }
                    if(!act) break; pair_iters++; pair_lane+=act; cost += C_PAIR; }
                { int lc=-1, cnt=0; for(int t=0;t<32;++t){ PL& p=ln[t]; if (p.pend>=0){ ++cnt; ++p.l.leaves; lc=std::max(lc,sphere(p.l,N[p.pend])); p.pend=-1; } }
                  if (lc>=0){ cost += (lc?C_LEAF1:C_LEAF0) + 4; leafphase++; leaflanes+=cnt; } }
                cost += 6;
            }
            for(int t=0;t<32;++t){ visits+=ln[t].l.visits; leaves+=ln[t].l.leaves; }
        }
        printf("%s pair scheme: pair-visits/ray=%.2f box-tests/ray=%.2f leaves/ray=%.2f warp-instr/ray=%.1f (at %d per pair iteration) iters/warp=%.1f lanes/iter=%.1f leafphases/warp=%.1f\n",
               rayfile, visits/nrays, boxtests/nrays, leaves/nrays, cost/nrays, C_PAIR, pair_iters/(nrays/32), pair_lane/pair_iters, leafphase/(nrays/32));
        return 0;
    }
    if (scheme == 4) {
        // while-while (K = 1) with STRAGGLER EVICTION: at a phase boundary (all lanes at a leaf or done) a warp with at
        // most T lanes still walking writes them to a continuation queue and ends; the stragglers of an octant are
        // regrouped into dense warps and finished by a second pass (no eviction there; argv[6] = 1: evict again).
        const int C_EVICT = 24, C_RESUME = 30; const int again = argc > 6 ? atoi(argv[6]) : 0;
        double evicted = 0, evictions = 0;
        for (int oc=0;oc<8;++oc) {
            const std::vector<Node>& N = L[oc];
            std::vector<Lane> pool, nextpool;
            for (size_t w=0; w+32<=q[oc].size(); w+=32) for (int t=0;t<32;++t){ Lane l; l.r=q[oc][w+t]; for(int k=0;k<3;++k) l.inv[k]=1.0f/l.r.d[k];
                l.a=l.r.d[0]*l.r.d[0]+l.r.d[1]*l.r.d[1]+l.r.d[2]*l.r.d[2]; l.i=0; l.done=false; l.best=INFINITY; l.obj=-1; l.np=0; l.visits=l.leaves=0; pool.push_back(l); nrays+=1; }
            for (int pass=0; !pool.empty(); ++pass) {
                const bool may_evict = pass == 0 || (again && pass < 4);
                for (size_t w=0; w<pool.size(); w+=32) {
                    const int nl = (int)std::min<size_t>(32, pool.size()-w); Lane* ln=&pool[w];
                    if (pass) cost += C_RESUME;
                    for(;;){
                        int alive=0; for(int t=0;t<nl;++t) if(!ln[t].done) ++alive; if(!alive) break;
                        if (may_evict && alive <= T && alive < nl) { for(int t=0;t<nl;++t) if(!ln[t].done){ nextpool.push_back(ln[t]); ln[t].done=true; ++evicted; } cost += C_EVICT; ++evictions; break; }
                        outer++;
                        for(;;){ int act=0; for (int t=0;t<nl;++t){ Lane& l=ln[t]; if (l.done||l.np>=1) continue; ++act; const Node& n=N[l.i]; uint32_t m=fb(n.f[3]);
                                if (m < (1u<<30)) { ++l.visits; l.i = slab_miss(n,l,l.best)? m : l.i+1; }
                                else if (m==0xffffffffu) { l.done=true; }
                                else { l.pend[l.np++]=l.i; l.i++; } }
                            if(!act) break; warp_iters++; lane_iters+=act; cost += C_NODE; }
                        { int lc=-1, cnt=0; for(int t=0;t<nl;++t){ Lane& l=ln[t]; if (l.np>0){ ++cnt; ++l.leaves; lc=std::max(lc,sphere(l,N[l.pend[0]])); l.np=0; } }
                          if (lc>=0){ cost += (lc?C_LEAF1:C_LEAF0) + 4; leafphase++; leaflanes+=cnt; } }
                        cost += 6 + 3;  // + the vote
                    }
                }
                // a lane copied to nextpool keeps its visit counters; count the finished ones here
                for (auto& l: pool) { bool moved=false; (void)moved; }
                std::vector<Lane> fin; fin.swap(pool);
                for (auto& l: fin) { visits += 0; (void)l; }
                // visits/leaves: lanes evicted were copied WITH their counters, so only count lanes that were not copied
                // (a copied lane was marked done right after the copy; its counters are counted when its copy finishes)
                pool.swap(nextpool); nextpool.clear();
                (void)fin;
            }
        }
        printf("evicted %.1f %% of the rays in %.0f evictions\n", 100.0*evicted/nrays, evictions);
    }
    if (scheme != 2 && scheme != 3 && scheme != 4) for (int oc=0;oc<8;++oc) for (size_t w=0; w+32<=q[oc].size(); w+=32) {
        Lane ln[32];
        for (int t=0;t<32;++t){ Lane& l=ln[t]; l.r=q[oc][w+t]; for(int k=0;k<3;++k) l.inv[k]=1.0f/l.r.d[k];
            l.a=l.r.d[0]*l.r.d[0]+l.r.d[1]*l.r.d[1]+l.r.d[2]*l.r.d[2]; l.i=0; l.done=false; l.best=INFINITY; l.obj=-1; l.np=0; l.visits=l.leaves=0; }
        nrays+=32;
        const std::vector<Node>& N = L[oc];
        if (scheme == 0) { // plain if/else loop: one node per iteration per lane, leaf executed in same iteration
            for(;;){ bool anyI=false, anyL=false; int lc=0; int act=0;
                for (int t=0;t<32;++t){ Lane& l=ln[t]; if (l.done) continue; ++act; const Node& n=N[l.i]; uint32_t m=fb(n.f[3]);
                    if (m < (1u<<30)) { anyI=true; ++l.visits; l.i = slab_miss(n,l,l.best)? m : l.i+1; }
                    else if (m==0xffffffffu) { l.done=true; }
                    else { anyL=true; ++l.leaves; lc=std::max(lc,sphere(l,n)); l.i++; } }
                if (!act) break; warp_iters++; lane_iters+=act;
                cost += 3 + (anyI? C_NODE-3:0) + (anyL? (lc?C_LEAF1:C_LEAF0):0); }
        } else { // while-while with K pending leaves per lane
            for(;;){
                bool any=false; for(int t=0;t<32;++t) if(!ln[t].done) any=true; if(!any) break;
                outer++;
                // inner: walk until every live lane has K pending or is done
                for(;;){ int act=0; for (int t=0;t<32;++t){ Lane& l=ln[t]; if (l.done||l.np>=K) continue; ++act; const Node& n=N[l.i]; uint32_t m=fb(n.f[3]);
                        if (m < (1u<<30)) { ++l.visits; l.i = slab_miss(n,l,l.best)? m : l.i+1; }
                        else if (m==0xffffffffu) { l.done=true; }
                        else { l.pend[l.np++]=l.i; l.i++; } }
                    if(!act) break; warp_iters++; lane_iters+=act; cost += C_NODE + (K>1?1:0); }
                for (int j=0;j<K;++j){ int lc=-1, cnt=0; for(int t=0;t<32;++t){ Lane& l=ln[t]; if (l.np>j){ ++cnt; ++l.leaves; lc=std::max(lc,sphere(l,N[l.pend[j]])); } }
                    if (lc>=0){ cost += (lc?C_LEAF1:C_LEAF0) + 4; leafphase++; leaflanes+=cnt; } }
                for(int t=0;t<32;++t) ln[t].np=0;
                cost += 6;
            }
        }
        for(int t=0;t<32;++t){ visits+=ln[t].visits; leaves+=ln[t].leaves; }
    }
    printf("%s scheme=%d K=%d rays=%.0f visits/ray=%.2f leaves/ray=%.2f warp-instr/ray=%.1f node-iters/warp=%.1f lanes/iter=%.1f outer/warp=%.1f leafphases/warp=%.1f lanes/leafphase=%.1f\n",
        rayfile, scheme, K, nrays, visits/nrays, leaves/nrays, cost/nrays, warp_iters/(nrays/32), lane_iters/warp_iters, outer/(nrays/32), leafphase/(nrays/32), leafphase? leaflanes/leafphase:0);
}
