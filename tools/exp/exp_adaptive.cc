// Experiment: adaptive-radius event graph + per-component dense blossom. Stats only.
#include "../../oracle/chromobius_oracle.cc"
#include <fstream>
#include <random>
#include <chrono>
using namespace orc;

struct Exp {
    Tables &t;
    std::vector<int64_t> dist; std::vector<uint32_t> touched; std::vector<int32_t> evidx;
    uint64_t node_visits = 0; size_t max_touched=0,max_found=0; long touched_hist[64]={0}, found_hist[32]={0};
    Exp(Tables &tt) : t(tt) { dist.assign(t.num_dets*2, -1); evidx.assign(t.num_dets*2, -1); }
    // explore from src up to radius R (inclusive); if R<0: until first event found, then finish that distance level.
    // returns list (event idx, d) sorted by d; sets *outR. min_new: need at least one event not in `have` count (explore until count of found > have)
    void explore(node_t src, int64_t R, size_t need_more_than, std::vector<std::pair<int,int64_t>> &found, int64_t *outR) {
        typedef std::pair<int64_t,node_t> QE;
        std::priority_queue<QE,std::vector<QE>,std::greater<QE>> pq;
        touched.clear(); found.clear();
        dist[src]=0; touched.push_back(src); pq.push({0,src});
        int64_t limit = R;  // -1 = unknown yet
        while(!pq.empty()){
            auto [d,u]=pq.top(); pq.pop();
            if(d!=dist[u]) continue;
            if (limit>=0 && d>limit) break;
            node_visits++;
            if(u!=src && evidx[u]>=0){ found.push_back({evidx[u],d}); if(limit<0 && found.size()>need_more_than) limit=d; }
            for(uint32_t k=t.row_ptr[u];k<t.row_ptr[u+1];k++){ node_t v=t.nbr[k]; int64_t nd=d+t.wt[k];
                if(limit>=0 && nd>limit) continue;
                if(dist[v]<0||nd<dist[v]){ if(dist[v]<0) touched.push_back(v); dist[v]=nd; pq.push({nd,v}); } }
        }
        max_touched=std::max<size_t>(max_touched,touched.size()); touched_hist[std::min<size_t>(63, touched.size()/8)]++; max_found=std::max<size_t>(max_found,found.size()); found_hist[std::min<size_t>(31,found.size())]++;
        for(auto v:touched) dist[v]=-1;
        *outR = limit;
    }
};

int main(int argc,char**argv){
    std::ifstream f(argv[1]); std::stringstream ss; ss<<f.rdbuf();
    auto t0=std::chrono::steady_clock::now();
    auto t = compile(ss.str());
    auto t1=std::chrono::steady_clock::now();
    fprintf(stderr,"compiled D=%lu in %.2fs\n",(unsigned long)t->num_dets,std::chrono::duration<double>(t1-t0).count());
    // flat errors for sampling
    Dem dem = Dem::parse(ss.str()); std::vector<FlatError> errs; {uint64_t off=0,md=0,mo=0;bool a=false,b=false; flatten_block(dem.instructions,off,&errs,&md,&mo,&a,&b);}
    int shots = argc>2?atoi(argv[2]):100;
    std::mt19937_64 rng(1);
    Exp ex(*t); bool check = argc>3; long mism=0; Decoder chk(std::make_unique<Tables>(*t));
    std::map<int,long> comp_hist; long rounds_total=0, solves=0; long max_comp=0; double sum_n3=0; long total_events=0;
    std::vector<long> round_hist(50,0);
    for(int s=0;s<shots;s++){
        std::vector<uint8_t> bits(t->num_dets,0);
        for(auto&e:errs){ // slow but fine
            if(std::generate_canonical<double,53>(rng)<e.p) for(auto&tg:e.targets) if(tg.kind==0) bits[tg.val]^=1; }
        std::vector<uint64_t> ev; for(size_t d=0;d<t->num_dets;d++) if(bits[d]&&!t->node_colors[d].ignored){ev.push_back(2*d);ev.push_back(2*d+1);}
        int m=ev.size(); total_events+=m; if(!m) continue;
        for(int i=0;i<m;i++) ex.evidx[ev[i]]=i;
        std::vector<int64_t> R(m); std::vector<std::vector<std::pair<int,int64_t>>> nb(m);
        for(int i=0;i<m;i++) ex.explore(ev[i],-1,0,nb[i],&R[i]);
        std::vector<char> dirty_ev(m,1);
        std::vector<int> partner(m,-1);
        int round=0;
        while(true){
            round++;
            // union find
            std::vector<int> uf(m); for(int i=0;i<m;i++) uf[i]=i;
            std::function<int(int)> find=[&](int x){while(uf[x]!=x){uf[x]=uf[uf[x]];x=uf[x];}return x;};
            for(int i=0;i<m;i++) for(auto&p:nb[i]){int a=find(i),b=find(p.first); if(a!=b) uf[a]=b;}
            std::map<int,std::vector<int>> comps; for(int i=0;i<m;i++) comps[find(i)].push_back(i);
            std::vector<std::pair<int,int64_t>> grow; // (event, newR or -1 for next-event)
            for(auto&[root,mem]:comps){
                bool dirty=false; for(int i:mem) dirty|=dirty_ev[i];
                if(!dirty) continue;
                int n=mem.size();
                if(round==1 || true) {}
                if(n%2){ for(int i:mem) grow.push_back({i,-1}); continue; }
                std::map<int,int> loc; for(int k=0;k<n;k++) loc[mem[k]]=k+1;
                DenseBlossom bl; bl.init(n); int64_t big=2; for(int i:mem) for(auto&p:nb[i]) big=std::max(big,p.second+2);
                for(int i:mem) for(auto&p:nb[i]) bl.add_edge(loc[i],loc[p.first],big-p.second);
                solves++; sum_n3+= (double)n*n*n; comp_hist[n]++; max_comp=std::max<long>(max_comp,n);
                int cnt=bl.solve();
                if(cnt*2!=n){ for(int i:mem) if(!bl.match[loc[i]]) grow.push_back({i,-1}); 
                    // also grow all (safer)
                    continue; }
                bool okc=true;
                for(int i:mem){ int64_t need=big-bl.lab[loc[i]]; if(need>R[i]){ grow.push_back({i,need}); okc=false; } }
                if(okc) for(int i:mem){ partner[i]=mem[bl.match[loc[i]]-1]; int j=partner[i]; bool has=false; for(auto&p:nb[i]) has|=p.first==j; for(auto&p:nb[j]) has|=p.first==i; if(!has){ fprintf(stderr,"BAD ASSIGN shot %d round %d n %d i %d j %d loc %d %d w=%ld n_x=%d st[i]=%d\n",s,round,n,i,j,loc[i],loc[j],(long)bl.g[loc[i]][loc[j]].w, bl.n_x, bl.st[loc[i]]); } }
            }
            std::fill(dirty_ev.begin(),dirty_ev.end(),0);
            if(grow.empty()) break;
            for(auto&[i,nr]:grow){
                std::vector<std::pair<int,int64_t>> fnd; int64_t outR;
                if(nr<0) ex.explore(ev[i],-1,nb[i].size(),fnd,&outR); else ex.explore(ev[i],nr,0,fnd,&outR);
                if(outR<0){ fprintf(stderr,"no more events reachable shot %d\n",s); return 1; }
                nb[i]=fnd; R[i]=outR; dirty_ev[i]=1;
                for(auto&p:fnd) dirty_ev[p.first]=1;
            }
            // mark comps containing dirty events: handled by dirty check (members). but comps merged: new comp contains dirty ev -> ok
            if(round>40){fprintf(stderr,"too many rounds\n");break;}
        }
        for(int i=0;i<m;i++) ex.evidx[ev[i]]=-1;
        { int64_t w=0; for(int i=0;i<m;i++){ int j=partner[i]; if(j<0){fprintf(stderr,"unmatched!\n");return 1;} if(j>i){ int64_t d=-1; for(auto&p:nb[i]) if(p.first==j) d=p.second; for(auto&p:nb[j]) if(p.first==i) d=p.second; if(d<0){ for(int q:{i,j}){fprintf(stderr,"ev %d R=%ld nb:",q,(long)R[q]); for(auto&p:nb[q]) fprintf(stderr," (%d,%ld)",p.first,(long)p.second); fprintf(stderr,"\n");} fprintf(stderr,"missing d shot %d i %d j %d partner[j] %d m %d\n",s,i,j,partner[j],m);return 1;} w+=d; } }
          if(check){ MatchResult mr; if(!chk.match_events(ev,&mr)){fprintf(stderr,"oracle fail\n");return 1;} if(mr.weight!=w){ fprintf(stderr,"WEIGHT MISMATCH shot %d: %ld vs oracle %ld\n",s,(long)w,(long)mr.weight); mism++; } } }
        rounds_total+=round; round_hist[std::min(round,49)]++;
        for(int i=0;i<m;i++) ex.evidx[ev[i]]=-1;
    }
    printf("shots %d events/shot %.1f  rounds/shot %.2f solves/shot %.1f max_comp %ld sum_n3/shot %.0f node_visits/shot %.0f\n",shots,(double)total_events/shots,(double)rounds_total/shots,(double)solves/shots,max_comp,sum_n3/shots,(double)ex.node_visits/shots);
    printf("mismatches %ld (check=%d)\n",mism,(int)check);
    printf("max_touched %zu max_found %zu\n touched/8 hist:",ex.max_touched,ex.max_found); for(int i=0;i<64;i++) if(ex.touched_hist[i]) printf(" %d:%ld",i,ex.touched_hist[i]); printf("\n found hist:"); for(int i=0;i<32;i++) if(ex.found_hist[i]) printf(" %d:%ld",i,ex.found_hist[i]); printf("\n");
    printf("round hist:"); for(int i=0;i<50;i++) if(round_hist[i]) printf(" %d:%ld",i,round_hist[i]); printf("\n");
    printf("comp size hist (size:count):"); for(auto&[k,v]:comp_hist) printf(" %d:%ld",k,v); printf("\n");
}
