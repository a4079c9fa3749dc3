// ============================================================================
// oracle/cityflow_oracle.cpp  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// A single-threaded, fp64, object-per-vehicle CPU restatement of the CityFlow
// microscopic engine that rbokade/pytsc drives through `cityflow.Engine`
// (reference call sites: pytsc/backends/cityflow/simulator.py:71-77,86-88,95;
// retriever.py:35,95-97,109-111; traffic_signal.py:31,58).
//
// CityFlow itself is a third-party C++/pybind11 module that is NOT vendored
// under /root/reference and is not pinned by the reference's setup.py
// (setup.py:7-10).  It is not installable in the build container, so this file
// restates CityFlow's *published algorithm* (cityflow-project/CityFlow,
// src/engine/engine.cpp, src/vehicle/{vehicle,router}.cpp,
// src/roadnet/{roadnet,trafficlight}.cpp, src/flow/flow.cpp) as summarised in
// SURVEY.md Appendix A, with the engine configuration pytsc always uses
// (interval 1.0, laneChange false, rlTrafficLight true, thread_num 1).
//
// PARITY UNPINNED for the vehicle dynamics: the reference ships no golden
// trajectories, replays or known-answer tests, and the real
// engine cannot be run here.  What *is* pinned is everything downstream of the
// engine: the reference's own Python Retriever / TrafficSignal / reward /
// action-mask / observation classes are executed, unmodified, on top of this
// engine (oracle/engine.py gives it the cityflow.Engine call surface) to
// produce the golden fixtures in tests/golden/.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
// reference legs may load this library.  The product path (pytsc_b200/) never
// does, and has no CPU fallback.
//
// Layout mirrors CityFlow's object model on purpose (lists of vehicle pointers
// per drivable, a two-phase buffered update) so that it is an *independent*
// formulation from the data-parallel CUDA kernels it checks.
// ============================================================================
#include <algorithm>
#include <cassert>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <fstream>
#include <list>
#include <map>
#include <memory>
#include <random>
#include <set>
#include <sstream>
#include <string>
#include <unordered_set>
#include <vector>

namespace cfo {

// ----------------------------------------------------------------------------
// Minimal JSON reader (objects, arrays, numbers, strings, true/false/null).
// ----------------------------------------------------------------------------
struct Json {
    enum Kind { Null, Bool, Num, Str, Arr, Obj } kind = Null;
    bool b = false;
    double num = 0;
    std::string str;
    std::vector<Json> arr;
    std::vector<std::pair<std::string, Json>> obj;

    const Json *find(const std::string &k) const {
        for (auto &kv : obj) if (kv.first == k) return &kv.second;
        return nullptr;
    }
    const Json &at(const std::string &k) const {
        const Json *j = find(k);
        if (!j) throw std::runtime_error("json: missing key '" + k + "'");
        return *j;
    }
    double number(const std::string &k, double dflt) const {
        const Json *j = find(k);
        return (j && j->kind == Num) ? j->num : dflt;
    }
    bool boolean(const std::string &k, bool dflt) const {
        const Json *j = find(k);
        if (!j) return dflt;
        if (j->kind == Bool) return j->b;
        if (j->kind == Num) return j->num != 0;
        return dflt;
    }
};

struct JsonParser {
    const char *p, *end;
    explicit JsonParser(const std::string &s) : p(s.data()), end(s.data() + s.size()) {}
    void ws() { while (p < end && (*p == ' ' || *p == '\n' || *p == '\t' || *p == '\r')) ++p; }
    [[noreturn]] void fail(const char *m) { throw std::runtime_error(std::string("json parse error: ") + m); }
    Json parse() { ws(); Json j = value(); ws(); return j; }
    Json value() {
        ws();
        if (p >= end) fail("eof");
        Json j;
        char c = *p;
        if (c == '{') {
            j.kind = Json::Obj; ++p; ws();
            if (*p == '}') { ++p; return j; }
            while (true) {
                ws(); if (*p != '"') fail("key");
                std::string k = string();
                ws(); if (*p != ':') fail("colon"); ++p;
                j.obj.emplace_back(std::move(k), value());
                ws();
                if (*p == ',') { ++p; continue; }
                if (*p == '}') { ++p; break; }
                fail("object");
            }
        } else if (c == '[') {
            j.kind = Json::Arr; ++p; ws();
            if (*p == ']') { ++p; return j; }
            while (true) {
                j.arr.push_back(value());
                ws();
                if (*p == ',') { ++p; continue; }
                if (*p == ']') { ++p; break; }
                fail("array");
            }
        } else if (c == '"') {
            j.kind = Json::Str; j.str = string();
        } else if (c == 't') { j.kind = Json::Bool; j.b = true; p += 4; }
        else if (c == 'f') { j.kind = Json::Bool; j.b = false; p += 5; }
        else if (c == 'n') { j.kind = Json::Null; p += 4; }
        else {
            char *e = nullptr;
            j.kind = Json::Num; j.num = std::strtod(p, &e);
            if (e == p) fail("number");
            p = e;
        }
        return j;
    }
    std::string string() {
        std::string s; ++p;
        while (p < end && *p != '"') {
            if (*p == '\\') {
                ++p;
                switch (*p) {
                    case 'n': s += '\n'; break; case 't': s += '\t'; break;
                    case 'r': s += '\r'; break; case 'b': s += '\b'; break;
                    case 'f': s += '\f'; break;
                    case 'u': p += 4; s += '?'; break;
                    default: s += *p;
                }
                ++p;
            } else s += *p++;
        }
        ++p;
        return s;
    }
};

static Json readJsonFile(const std::string &path) {
    std::ifstream f(path);
    if (!f) throw std::runtime_error("cannot open " + path);
    std::stringstream ss; ss << f.rdbuf();
    std::string s = ss.str();
    return JsonParser(s).parse();
}

// ----------------------------------------------------------------------------
// Geometry helpers (CityFlow utility.h semantics).
// ----------------------------------------------------------------------------
static constexpr double EPS = 1e-8;

struct Point { double x = 0, y = 0; };
static inline Point operator+(Point a, Point b) { return {a.x + b.x, a.y + b.y}; }
static inline Point operator-(Point a, Point b) { return {a.x - b.x, a.y - b.y}; }
static inline Point operator*(Point a, double k) { return {a.x * k, a.y * k}; }
static inline double plen(Point a) { return std::sqrt(a.x * a.x + a.y * a.y); }
static inline Point unit(Point a) { double l = plen(a); return {a.x / l, a.y / l}; }
static inline Point normal(Point a) { return {-a.y, a.x}; }
static inline double crossMul(Point a, Point b) { return a.x * b.y - a.y * b.x; }
static inline double dotMul(Point a, Point b) { return a.x * b.x + a.y * b.y; }
static inline int sgn(double x) { return (x + EPS > 0) - (x < EPS); }
static inline double min2double(double x, double y) { return x < y ? x : y; }
static inline double max2double(double x, double y) { return x > y ? x : y; }

static double polylineLength(const std::vector<Point> &pts) {
    double l = 0;
    for (size_t i = 0; i + 1 < pts.size(); ++i) l += plen(pts[i + 1] - pts[i]);
    return l;
}
static Point calcIntersectPoint(Point A, Point B, Point C, Point D) {
    Point P = A;
    double t = crossMul(A - C, C - D) / crossMul(A - B, C - D);
    P.x += (B.x - A.x) * t;
    P.y += (B.y - A.y) * t;
    return P;
}
static bool onSegment(Point A, Point B, Point P) {
    double v1 = crossMul(B - A, P - A);
    double v2 = dotMul(P - A, P - B);
    return sgn(v1) == 0 && sgn(v2) <= 0;
}

// ----------------------------------------------------------------------------
// Road network.
// ----------------------------------------------------------------------------
struct Vehicle; struct Road; struct Lane; struct LaneLink; struct RoadLink; struct Intersection; struct Cross;

enum RoadLinkType { go_straight = 3, turn_left = 2, turn_right = 1 };

struct Drivable {
    bool lane = true;
    int index = -1;           // lanes: 0..L-1 (road order, lane order); lane-links: L..L+K-1
    double length = 0;
    double maxSpeed = 0;
    std::list<Vehicle *> vehicles;   // front = first (most advanced) vehicle
    bool isLane() const { return lane; }
    bool isLaneLink() const { return !lane; }
    Vehicle *getFirstVehicle() const { return vehicles.empty() ? nullptr : vehicles.front(); }
    Vehicle *getLastVehicle() const { return vehicles.empty() ? nullptr : vehicles.back(); }
    virtual ~Drivable() {}
};

struct Lane : Drivable {
    Road *road = nullptr;
    int laneIndex = 0;
    double width = 0;
    std::vector<LaneLink *> laneLinks;
    std::deque<Vehicle *> waitingBuffer;
    std::string id;
    std::vector<LaneLink *> getLaneLinksToRoad(const Road *r) const;
    bool available(const Vehicle *v) const;
    bool canEnter(const Vehicle *v) const;
};

struct LaneLink : Drivable {
    RoadLink *roadLink = nullptr;
    Lane *startLane = nullptr, *endLane = nullptr;
    std::vector<Point> points;
    std::vector<Cross *> crosses;
    bool isAvailable() const;
    bool isTurn() const;
    RoadLinkType type() const;
};

struct RoadLink {
    RoadLinkType type = go_straight;
    Road *startRoad = nullptr, *endRoad = nullptr;
    Intersection *intersection = nullptr;
    int index = 0;
    std::vector<std::unique_ptr<LaneLink>> laneLinks;
};

struct Cross {
    LaneLink *laneLinks[2] = {nullptr, nullptr};
    Vehicle *notifyVehicles[2] = {nullptr, nullptr};
    double notifyDistances[2] = {0, 0};
    double distanceOnLane[2] = {0, 0};
    double leaveDistance = 0, arriveDistance = 30;
    void clearNotify() { notifyVehicles[0] = notifyVehicles[1] = nullptr; }
    double getDistanceByLane(const LaneLink *ll) const { return ll == laneLinks[0] ? distanceOnLane[0] : distanceOnLane[1]; }
    void notify(LaneLink *ll, Vehicle *v, double d) {
        int i = (ll == laneLinks[0]) ? 0 : 1;
        notifyVehicles[i] = v; notifyDistances[i] = d;
    }
    Vehicle *getFoeVehicle(const LaneLink *ll) const { return ll == laneLinks[0] ? notifyVehicles[1] : notifyVehicles[0]; }
    bool canPass(const Vehicle *vehicle, const LaneLink *laneLink, double distanceToLaneLinkStart) const;
};

struct LightPhase { double time = 0; std::vector<char> roadLinkAvailable; };

struct Intersection {
    std::string id;
    bool isVirtual = false;
    double width = 0;
    Point point;
    std::vector<Road *> roads;
    std::vector<std::unique_ptr<RoadLink>> roadLinks;
    std::vector<Cross> crosses;
    std::vector<LaneLink *> laneLinks;
    std::vector<LightPhase> phases;
    int curPhase = 0;
    void initCrosses();
};

struct Road {
    std::string id;
    Intersection *startIntersection = nullptr, *endIntersection = nullptr;
    std::vector<Point> points;
    std::vector<std::unique_ptr<Lane>> lanes;
    std::vector<Vehicle *> planRouteBuffer;
    void initLanesPoints();
};

std::vector<LaneLink *> Lane::getLaneLinksToRoad(const Road *r) const {
    std::vector<LaneLink *> ret;
    for (auto *ll : laneLinks) if (ll->endLane->road == r) ret.push_back(ll);
    return ret;
}
bool LaneLink::isAvailable() const {
    const Intersection *it = roadLink->intersection;
    return it->phases[it->curPhase].roadLinkAvailable[roadLink->index] != 0;
}
bool LaneLink::isTurn() const { return roadLink->type == turn_left || roadLink->type == turn_right; }
RoadLinkType LaneLink::type() const { return roadLink->type; }

// Trim both ends by the (non-virtual) intersection width; lane length is the
// length of the trimmed centre polyline offset sideways (A.1).
void Road::initLanesPoints() {
    double dsum = 0.0;
    std::vector<Point> rp = points;
    if (!startIntersection->isVirtual) {
        double w = startIntersection->width;
        Point p1 = rp[0], p2 = rp[1];
        rp[0] = p1 + unit(p2 - p1) * w;
    }
    if (!endIntersection->isVirtual) {
        double w = endIntersection->width;
        Point p1 = rp[rp.size() - 2], p2 = rp[rp.size() - 1];
        rp[rp.size() - 1] = p2 - unit(p2 - p1) * w;
    }
    for (auto &lp : lanes) {
        Lane &lane = *lp;
        double dmin = dsum, dmax = dsum + lane.width;
        std::vector<Point> lpts;
        int n = (int) rp.size();
        for (int j = 0; j < n; ++j) {
            Point u;
            if (j == 0) u = unit(rp[1] - rp[0]);
            else if (j + 1 == n) u = unit(rp[j] - rp[j - 1]);
            else {
                Point u1 = unit(rp[j + 1] - rp[j]), u2 = unit(rp[j] - rp[j - 1]);
                u = unit(u1 + u2);
            }
            Point v = normal(u) * -1.0;
            lpts.push_back(rp[j] + v * ((dmin + dmax) / 2.0));
        }
        lane.length = polylineLength(lpts);
        dsum += lane.width;
    }
}

// All pairwise polyline intersections of lane-links inside one intersection;
// first hit per pair (A.5).  Lane-links are visited road-link major.
void Intersection::initCrosses() {
    std::vector<LaneLink *> all;
    for (auto &rl : roadLinks) for (auto &ll : rl->laneLinks) all.push_back(ll.get());
    int n = (int) all.size();
    for (int i = 0; i < n; ++i) {
        for (int j = i + 1; j < n; ++j) {
            LaneLink *la = all[i], *lb = all[j];
            if (la->points.empty() || lb->points.empty()) continue;
            double disa = 0.0;
            bool found = false;
            for (size_t ia = 0; ia + 1 < la->points.size() && !found; ++ia) {
                double disb = 0.0;
                for (size_t ib = 0; ib + 1 < lb->points.size(); ++ib) {
                    Point A1 = la->points[ia], A2 = la->points[ia + 1];
                    Point B1 = lb->points[ib], B2 = lb->points[ib + 1];
                    if (sgn(crossMul(A2 - A1, B2 - B1)) == 0) { disb += plen(B2 - B1); continue; }
                    Point P = calcIntersectPoint(A1, A2, B1, B2);
                    if (onSegment(A1, A2, P) && onSegment(B1, B2, P)) {
                        Cross c;
                        c.laneLinks[0] = la; c.laneLinks[1] = lb;
                        c.distanceOnLane[0] = disa + plen(P - A1);
                        c.distanceOnLane[1] = disb + plen(P - B1);
                        crosses.push_back(c);
                        found = true;
                        break;
                    }
                    disb += plen(B2 - B1);
                }
                disa += plen(la->points[ia + 1] - la->points[ia]);
            }
        }
    }
    for (Cross &c : crosses) {
        c.laneLinks[0]->crosses.push_back(&c);
        c.laneLinks[1]->crosses.push_back(&c);
    }
    for (LaneLink *ll : all) {
        std::stable_sort(ll->crosses.begin(), ll->crosses.end(), [ll](Cross *ca, Cross *cb) {
            return ca->getDistanceByLane(ll) < cb->getDistanceByLane(ll);
        });
    }
}

// ----------------------------------------------------------------------------
// Vehicles.
// ----------------------------------------------------------------------------
struct VehicleInfo {
    double speed = 0;
    double len = 5, width = 2;
    double maxPosAcc = 4.5, maxNegAcc = 4.5, usualPosAcc = 2.5, usualNegAcc = 2.5;
    double minGap = 2, maxSpeed = 16.66667, headwayTime = 1;
    double yieldDistance = 5, turnSpeed = 8.3333;
};

struct Engine;
struct Flow;

struct Router {
    Vehicle *vehicle = nullptr;
    std::vector<Road *> route;
    size_t iCurRoad = 0;
    std::mt19937 *rnd = nullptr;
    mutable std::deque<Drivable *> planned;

    static LaneLink *selectLaneLink(const Lane *cur, const std::vector<LaneLink *> &lls) {
        if (lls.empty()) return nullptr;
        int best = 0; LaneLink *sel = nullptr;
        for (auto *ll : lls) {
            int d = std::abs(ll->endLane->laneIndex - cur->laneIndex);
            if (sel == nullptr || d < best) { best = d; sel = ll; }
        }
        return sel;
    }
    Drivable *getFirstDrivable() const {
        const auto &lanes = route[0]->lanes;
        std::vector<Lane *> cand;
        if (route.size() == 1) { for (auto &l : lanes) cand.push_back(l.get()); }
        else for (auto &l : lanes) if (!l->getLaneLinksToRoad(route[1]).empty()) cand.push_back(l.get());
        assert(!cand.empty());
        return cand[(*rnd)() % cand.size()];
    }
    Drivable *nextOf(const Drivable *cur) const {
        if (cur->isLaneLink()) return static_cast<const LaneLink *>(cur)->endLane;
        const Lane *cl = static_cast<const Lane *>(cur);
        size_t t = iCurRoad;
        while (t < route.size() && route[t] != cl->road) ++t;
        assert(t < route.size());
        if (t + 1 == route.size()) return nullptr;
        std::vector<LaneLink *> lls = cl->getLaneLinksToRoad(route[t + 1]);
        if (t + 2 == route.size()) return selectLaneLink(cl, lls);
        std::vector<LaneLink *> cand;
        for (auto *ll : lls) if (!ll->endLane->getLaneLinksToRoad(route[t + 2]).empty()) cand.push_back(ll);
        return selectLaneLink(cl, cand);
    }
    Drivable *getNextDrivable(size_t i) const;
    void update();
    bool isRouteValid() const {
        for (size_t i = 0; i + 1 < route.size(); ++i) {
            bool ok = false;
            for (auto &l : route[i]->lanes) if (!l->getLaneLinksToRoad(route[i + 1]).empty()) ok = true;
            if (!ok) return false;
        }
        return !route.empty();
    }
};

struct Vehicle {
    VehicleInfo info;
    int uid = -1;             // creation order
    int flowIndex = -1, flowCnt = 0;
    int priority = 0;
    double enterTime = 0;
    Engine *engine = nullptr;
    Flow *flow = nullptr;
    // controller info
    double dis = 0;
    Drivable *drivable = nullptr, *prevDrivable = nullptr;
    double approachingIntersectionDistance = 0;
    double gap = 0;
    int enterLaneLinkTime = INT_MAX;
    Vehicle *leader = nullptr, *blocker = nullptr;
    bool end = false, running = false;
    Router router;
    // two-phase buffer
    struct Buffer {
        bool isDisSet = false, isSpeedSet = false, isDrivableSet = false, isEndSet = false,
             isEnterLaneLinkTimeSet = false, isBlockerSet = false;
        double dis = 0, deltaDis = 0, speed = 0;
        Drivable *drivable = nullptr;
        bool end = false;
        int enterLaneLinkTime = 0;
        Vehicle *blocker = nullptr;
    } buffer;

    Drivable *getNextDrivable(int i = 0) const { return router.getNextDrivable((size_t) i); }
    double getMinBrakeDistance() const { return 0.5 * info.speed * info.speed / info.maxNegAcc; }
    bool canYield(double dist) const {
        return (dist > 0 && getMinBrakeDistance() < dist - info.yieldDistance) || (dist < 0 && dist + info.len < 0);
    }
    void setBlocker(Vehicle *b) { buffer.blocker = b; buffer.isBlockerSet = true; }
    Drivable *getChangedDrivable() const { return buffer.isDrivableSet ? buffer.drivable : nullptr; }

    void updateLeaderAndGap(Vehicle *leaderOnSameDrivable);
    double getNoCollisionSpeed(double vL, double dL, double vF, double dF, double gap_, double interval, double targetGap) const;
    double getCarFollowSpeed(double interval) const;
    double getBrakeDistanceAfterAccel(double acc, double dec, double interval) const;
    double getStopBeforeSpeed(double distance, double interval) const;
    bool isIntersectionRelated() const;
    double getIntersectionRelatedSpeed(double interval);
    double getNextSpeed(double interval);
    double getDistanceUntilSpeed(double speed, double acc) const;
    int getReachSteps(double distance, double targetSpeed, double acc) const;
    int getReachStepsOnLaneLink(double distance, const LaneLink *ll) const;
    void setDeltaDistance(double d);
    void update();
};

Drivable *Router::getNextDrivable(size_t i) const {
    if (i < planned.size()) return planned[i];
    // CityFlow extends the plan one element at a time; callers always ask for
    // i <= planned.size().
    while (planned.size() <= i) {
        const Drivable *from = planned.empty() ? vehicle->drivable : planned.back();
        if (from == nullptr) { planned.push_back(nullptr); continue; }
        planned.push_back(nextOf(from));
    }
    return planned[i];
}
void Router::update() {
    const Drivable *cur = vehicle->drivable;
    if (cur->isLane()) {
        const Lane *cl = static_cast<const Lane *>(cur);
        while (iCurRoad < route.size() && cl->road != route[iCurRoad]) ++iCurRoad;
    }
    for (auto it = planned.begin(); it != planned.end();) {
        if (*it != cur) it = planned.erase(it);
        else { it = planned.erase(it); break; }
    }
}

bool Lane::available(const Vehicle *v) const {
    if (!vehicles.empty()) {
        Vehicle *tail = vehicles.back();
        return tail->dis > tail->info.len + v->info.minGap;
    }
    return true;
}
bool Lane::canEnter(const Vehicle *v) const {
    if (!vehicles.empty()) {
        Vehicle *tail = vehicles.back();
        return tail->dis > tail->info.len + v->info.len || tail->info.speed >= 2;
    }
    return true;
}

struct Flow {
    VehicleInfo tmpl;
    std::vector<Road *> route;
    double interval = 1, startTime = 0, endTime = -1;
    double nowTime = 0, currentTime = 0;
    int cnt = 0, index = 0;
    bool valid = true;
    void reset() { nowTime = interval; currentTime = 0; cnt = 0; }
};

// ----------------------------------------------------------------------------
// Engine.
// ----------------------------------------------------------------------------
struct Engine {
    std::vector<std::unique_ptr<Road>> roads;
    std::vector<std::unique_ptr<Intersection>> intersections;
    std::vector<Lane *> lanes;
    std::vector<LaneLink *> laneLinks;
    std::vector<Drivable *> drivables;
    std::map<std::string, Road *> roadMap;
    std::map<std::string, Intersection *> interMap;
    std::vector<Flow> flows;

    std::map<int, Vehicle *> vehiclePool;      // priority -> vehicle (all created, unfinished)
    std::vector<Vehicle *> creationOrder;      // uid -> vehicle (nullptr once finished)
    std::vector<std::string> vehicleNames;     // uid -> "flow_i_c"
    std::set<Vehicle *> vehicleRemoveBuffer;
    std::vector<std::pair<Vehicle *, double>> pushBuffer;

    std::mt19937 rnd;
    double interval = 1.0;
    int seed = 0;
    bool rlTrafficLight = true, laneChange = false;
    size_t step = 0;
    int activeVehicleCount = 0, finishedVehicleCnt = 0;
    double cumulativeTravelTime = 0;
    long long nonFifoEvents = 0;               // diagnostics: a non-front vehicle left its drivable
    std::string lastError;

    double getCurrentTime() const { return step * interval; }
    bool checkPriority(int p) const { return vehiclePool.find(p) != vehiclePool.end(); }

    void loadConfig(const std::string &cfgFile);
    void loadRoadNet(const std::string &path);
    void loadFlow(const std::string &path);

    void nextStep();
    void flowStep(Flow &f);
    void planRoute();
    void handleWaiting();
    void notifyCross();
    void getAction();
    void vehicleControl(Vehicle &v);
    void updateLocation();
    void updateAction();
    void updateLeaderAndGap();
    void reset(bool resetRnd);
    double getAverageTravelTime() const;
};

// -- Vehicle dynamics (A.4, A.5, A.7) -----------------------------------------
void Vehicle::updateLeaderAndGap(Vehicle *ld) {
    if (ld != nullptr && ld->drivable == drivable) {
        leader = ld;
        gap = ld->dis - ld->info.len - dis;
        return;
    }
    leader = nullptr;
    Drivable *d = nullptr;
    Vehicle *cand = nullptr;
    double candGap = 0;
    double dist = drivable->length - dis;
    for (int i = 0;; ++i) {
        d = getNextDrivable(i);
        if (d == nullptr) return;
        if (d->isLaneLink()) {
            // lane-links leaving the same lane overlap: look at all of them
            for (LaneLink *ll : static_cast<LaneLink *>(d)->startLane->laneLinks) {
                if ((cand = ll->getLastVehicle()) != nullptr) {
                    candGap = dist + cand->dis - cand->info.len;
                    if (leader == nullptr || candGap < gap) { leader = cand; gap = candGap; }
                }
            }
            if (leader) return;
        } else {
            if ((leader = d->getLastVehicle()) != nullptr) {
                gap = dist + leader->dis - leader->info.len;
                return;
            }
        }
        dist += d->length;
        if (dist > info.maxSpeed * info.maxSpeed / info.usualNegAcc / 2 + info.maxSpeed * engine->interval * 2) return;
    }
}

double Vehicle::getNoCollisionSpeed(double vL, double dL, double vF, double dF, double gap_, double interval, double targetGap) const {
    double c = vF * interval / 2 + targetGap - 0.5 * vL * vL / dL - gap_;
    double a = 0.5 / dF;
    double b = 0.5 * interval;
    if (b * b < 4 * a * c) return -100;
    double v1 = 0.5 / a * (std::sqrt(b * b - 4 * a * c) - b);
    double v2 = 2 * vL - dL * interval + 2 * (gap_ - targetGap) / interval;
    return min2double(v1, v2);
}

double Vehicle::getCarFollowSpeed(double interval) const {
    if (leader == nullptr) return info.maxSpeed;
    double v = getNoCollisionSpeed(leader->info.speed, leader->info.maxNegAcc, info.speed, info.maxNegAcc, gap, interval, 0);
    double assumeDecel = 0, leaderSpeed = leader->info.speed;
    if (info.speed > leaderSpeed) assumeDecel = info.speed - leaderSpeed;
    v = min2double(v, getNoCollisionSpeed(leader->info.speed, assumeDecel, info.speed, info.maxNegAcc, gap, interval, info.minGap));
    v = min2double(v, (gap + (leaderSpeed + assumeDecel / 2) * interval - info.speed * interval / 2) / (info.headwayTime + interval / 2));
    return v;
}

double Vehicle::getBrakeDistanceAfterAccel(double acc, double dec, double interval) const {
    double cur = info.speed;
    double nxt = cur + acc * interval;
    return (cur + nxt) * interval / 2 + (nxt * nxt / dec / 2);
}

// (int) of an out-of-range double is what x86 cvttsd2si gives the original
// binary: INT_MIN.  Written out so that the CUDA path can state the same rule.
static inline int truncToIntX86(double x) {
    if (!(x > -2147483649.0 && x < 2147483648.0)) return INT_MIN;
    return (int) x;
}

double Vehicle::getStopBeforeSpeed(double distance, double interval) const {
    if (getBrakeDistanceAfterAccel(info.usualPosAcc, info.usualNegAcc, interval) < distance)
        return info.speed + info.usualPosAcc * interval;
    double takeInterval = 2 * distance / (info.speed + EPS) / interval;
    if (takeInterval >= 1) return info.speed - info.speed / truncToIntX86(takeInterval);
    return info.speed - info.speed / takeInterval;
}

bool Vehicle::isIntersectionRelated() const {
    if (drivable->isLaneLink()) return true;
    Drivable *nd = getNextDrivable();
    return nd && nd->isLaneLink() && drivable->length - dis <= approachingIntersectionDistance;
}

double Vehicle::getDistanceUntilSpeed(double speed, double acc) const {
    if (speed <= info.speed) return 0;
    double interval = engine->interval;
    int stage1steps = (int) std::floor((speed - info.speed) / acc / interval);
    double stage1speed = info.speed + stage1steps * acc / interval;
    double stage1dis = (info.speed + stage1speed) * (stage1steps * interval) / 2;
    return stage1dis + (stage1speed < speed ? ((stage1speed + speed) * interval / 2) : 0);
}
int Vehicle::getReachSteps(double distance, double targetSpeed, double acc) const {
    if (distance <= 0) return -1;
    if (info.speed > targetSpeed) return (int) std::ceil(distance / info.speed);
    double dUntil = getDistanceUntilSpeed(targetSpeed, acc);
    double interval = engine->interval;
    if (dUntil > distance)
        return (int) std::ceil((std::sqrt(info.speed * info.speed + 2 * acc * distance) - info.speed) / acc / interval);
    return (int) std::ceil((targetSpeed - info.speed) / acc / interval) + (int) std::ceil((distance - dUntil) / targetSpeed / interval);
}
int Vehicle::getReachStepsOnLaneLink(double distance, const LaneLink *ll) const {
    return getReachSteps(distance, ll->isTurn() ? info.turnSpeed : info.maxSpeed, info.usualPosAcc);
}

bool Cross::canPass(const Vehicle *vehicle, const LaneLink *laneLink, double distanceToLaneLinkStart) const {
    int i = (laneLink == laneLinks[0]) ? 0 : 1;
    Vehicle *foe = notifyVehicles[1 - i];
    RoadLinkType t1 = laneLinks[i]->type(), t2 = laneLinks[1 - i]->type();
    double d1 = distanceOnLane[i] - distanceToLaneLinkStart;
    double d2 = notifyDistances[1 - i];
    if (foe == nullptr) return true;
    if (!vehicle->canYield(d1)) return true;
    int yield = 0;
    if (!foe->canYield(d2)) yield = 1;
    if (yield == 0) {
        if (t1 > t2) yield = -1;
        else if (t1 < t2) {
            if (d2 > 0) {
                int foeSteps = foe->getReachStepsOnLaneLink(d2, laneLinks[1 - i]);
                int mySteps = vehicle->getReachStepsOnLaneLink(d1, laneLinks[i]);
                if (foeSteps > mySteps) yield = -1;
            } else if (d2 + foe->info.len < 0) yield = -1;
            if (yield == 0) yield = 1;
        } else {
            if (d2 > 0) {
                int foeSteps = foe->getReachStepsOnLaneLink(d2, laneLinks[1 - i]);
                int mySteps = vehicle->getReachStepsOnLaneLink(d1, laneLinks[i]);
                if (foeSteps > mySteps) yield = -1;
                else if (foeSteps < mySteps) yield = 1;
                else if (vehicle->enterLaneLinkTime == foe->enterLaneLinkTime) {
                    if (d1 == d2) yield = vehicle->priority > foe->priority ? -1 : 1;
                    else yield = d1 < d2 ? -1 : 1;
                } else yield = vehicle->enterLaneLinkTime < foe->enterLaneLinkTime ? -1 : 1;
            } else yield = d2 + foe->info.len < 0 ? -1 : 1;
        }
    }
    if (yield == 1) {   // deadlock: the foe's blocker chain loops
        Vehicle *fast = foe, *slow = foe;
        while (fast != nullptr && fast->blocker != nullptr) {
            slow = slow->blocker;
            fast = fast->blocker->blocker;
            if (slow == fast) { yield = -1; break; }
        }
    }
    return yield == -1;
}

double Vehicle::getIntersectionRelatedSpeed(double interval) {
    double v = info.maxSpeed;
    Drivable *nd = getNextDrivable();
    const LaneLink *ll = nullptr;
    if (nd && nd->isLaneLink()) {
        ll = static_cast<const LaneLink *>(nd);
        if (!ll->isAvailable() || !ll->endLane->canEnter(this)) {
            if (getMinBrakeDistance() > drivable->length - dis) {
                // cannot stop before the line: keep going
            } else {
                v = min2double(v, getStopBeforeSpeed(drivable->length - dis, interval));
                return v;
            }
        }
        if (ll->isTurn()) v = min2double(v, info.turnSpeed);
    }
    if (ll == nullptr && drivable->isLaneLink()) ll = static_cast<const LaneLink *>(drivable);
    double distanceToLaneLinkStart = drivable->isLane() ? -(drivable->length - dis) : dis;
    for (Cross *cross : ll->crosses) {
        double dOn = cross->getDistanceByLane(ll);
        if (dOn < distanceToLaneLinkStart) continue;
        if (!cross->canPass(this, ll, distanceToLaneLinkStart)) {
            v = min2double(v, getStopBeforeSpeed(dOn - distanceToLaneLinkStart - info.yieldDistance, interval));
            setBlocker(cross->getFoeVehicle(ll));
            break;
        }
    }
    return v;
}

double Vehicle::getNextSpeed(double interval) {
    double v = info.maxSpeed;
    v = min2double(v, info.speed + info.maxPosAcc * interval);
    v = min2double(v, drivable->maxSpeed);
    v = min2double(v, getCarFollowSpeed(interval));
    if (isIntersectionRelated()) v = min2double(v, getIntersectionRelatedSpeed(interval));
    v = max2double(v, info.speed - info.maxNegAcc * interval);
    return v;
}

void Vehicle::setDeltaDistance(double d) {
    if (!buffer.isDisSet || d < buffer.deltaDis) {
        buffer.isEndSet = false; buffer.isDrivableSet = false;
        buffer.deltaDis = d;
        d = d + dis;
        Drivable *dr = drivable;
        for (int i = 0; dr && d > dr->length; ++i) {
            d -= dr->length;
            Drivable *nd = router.getNextDrivable((size_t) i);
            if (nd == nullptr) { buffer.end = true; buffer.isEndSet = true; }
            dr = nd;
            buffer.drivable = dr; buffer.isDrivableSet = true;
        }
        buffer.dis = d; buffer.isDisSet = true;
    }
}

void Vehicle::update() {
    if (buffer.isEndSet) { end = buffer.end; buffer.isEndSet = false; }
    if (buffer.isDisSet) { dis = buffer.dis; buffer.isDisSet = false; }
    if (buffer.isSpeedSet) { info.speed = buffer.speed; buffer.isSpeedSet = false; }
    if (buffer.isDrivableSet) {
        prevDrivable = drivable;
        drivable = buffer.drivable;
        buffer.isDrivableSet = false;
        router.update();
    }
    if (buffer.isEnterLaneLinkTimeSet) { enterLaneLinkTime = buffer.enterLaneLinkTime; buffer.isEnterLaneLinkTimeSet = false; }
    if (buffer.isBlockerSet) { blocker = buffer.blocker; buffer.isBlockerSet = false; }
    else blocker = nullptr;
}

// -- Engine phases (A.2) -------------------------------------------------------
void Engine::flowStep(Flow &f) {
    if (!f.valid) return;
    if (f.endTime != -1 && f.currentTime > f.endTime) return;
    if (f.currentTime >= f.startTime) {
        while (f.nowTime >= f.interval) {
            Vehicle *v = new Vehicle();
            v->info = f.tmpl;
            v->engine = this; v->flow = &f;
            v->flowIndex = f.index; v->flowCnt = f.cnt++;
            v->enterTime = getCurrentTime();
            v->approachingIntersectionDistance =
                v->info.maxSpeed * v->info.maxSpeed / v->info.usualNegAcc / 2 + v->info.maxSpeed * interval * 2;
            v->router.vehicle = v; v->router.route = f.route; v->router.rnd = &rnd;
            int pr = (int) rnd();
            while (checkPriority(pr)) pr = (int) rnd();
            v->priority = pr;
            v->uid = (int) creationOrder.size();
            creationOrder.push_back(v);
            vehicleNames.push_back("flow_" + std::to_string(f.index) + "_" + std::to_string(v->flowCnt));
            vehiclePool.emplace(pr, v);
            f.route[0]->planRouteBuffer.push_back(v);
            f.nowTime -= f.interval;
        }
        f.nowTime += interval;
    }
    f.currentTime += interval;
}

void Engine::planRoute() {
    for (auto &road : roads) {
        for (Vehicle *v : road->planRouteBuffer) {
            if (v->router.isRouteValid()) {
                v->drivable = v->router.getFirstDrivable();
                static_cast<Lane *>(v->drivable)->waitingBuffer.push_back(v);
            } else {
                if (v->flow) v->flow->valid = false;
                vehiclePool.erase(v->priority);
                creationOrder[v->uid] = nullptr;
                delete v;
            }
        }
        road->planRouteBuffer.clear();
    }
}

void Engine::handleWaiting() {
    for (Lane *lane : lanes) {
        auto &buf = lane->waitingBuffer;
        if (buf.empty()) continue;
        Vehicle *v = buf.front();
        if (lane->available(v)) {
            v->running = true;
            activeVehicleCount += 1;
            Vehicle *tail = lane->getLastVehicle();
            lane->vehicles.push_back(v);
            v->updateLeaderAndGap(tail);
            buf.pop_front();
        }
    }
}

void Engine::notifyCross() {
    for (auto &it : intersections) for (Cross &c : it->crosses) c.clearNotify();
    for (auto &it : intersections) {
        for (LaneLink *ll : it->laneLinks) {
            const auto &crosses = ll->crosses;
            auto rIter = crosses.rbegin();
            // the vehicle that has just left onto the end lane
            Vehicle *v = ll->endLane->getLastVehicle();
            if (v && v->prevDrivable == ll) {
                double vehDistance = v->dis - v->info.len;
                while (rIter != crosses.rend()) {
                    double crossDistance = ll->length - (*rIter)->getDistanceByLane(ll);
                    if (crossDistance + vehDistance < (*rIter)->leaveDistance) {
                        (*rIter)->notify(ll, v, -(v->dis + crossDistance));
                        ++rIter;
                    } else break;
                }
            }
            // vehicles on the lane-link, front to back
            for (Vehicle *lv : ll->vehicles) {
                double vehDistance = lv->dis;
                while (rIter != crosses.rend()) {
                    double crossDistance = (*rIter)->getDistanceByLane(ll);
                    if (vehDistance > crossDistance) {
                        if (vehDistance - crossDistance - lv->info.len <= (*rIter)->leaveDistance)
                            (*rIter)->notify(ll, lv, crossDistance - vehDistance);
                        else break;
                    } else {
                        (*rIter)->notify(ll, lv, crossDistance - vehDistance);
                    }
                    ++rIter;
                }
            }
            // the first vehicle on the incoming lane, if it is heading here on green
            v = ll->startLane->getFirstVehicle();
            if (v && v->getNextDrivable() == ll && ll->isAvailable()) {
                double vehDistance = ll->startLane->length - v->dis;
                while (rIter != crosses.rend()) {
                    (*rIter)->notify(ll, v, vehDistance + (*rIter)->getDistanceByLane(ll));
                    ++rIter;
                }
            }
        }
    }
}

void Engine::vehicleControl(Vehicle &v) {
    double nextSpeed = v.getNextSpeed(interval);
    double deltaDis, speed = v.info.speed;
    if (nextSpeed < 0) {
        deltaDis = 0.5 * speed * speed / v.info.maxNegAcc;
        nextSpeed = 0;
    } else {
        deltaDis = (speed + nextSpeed) * interval / 2;
    }
    v.buffer.speed = nextSpeed; v.buffer.isSpeedSet = true;
    v.setDeltaDistance(deltaDis);
    if (!v.buffer.isEndSet && v.buffer.isDrivableSet) pushBuffer.emplace_back(&v, v.buffer.dis);
}

void Engine::getAction() {
    for (Vehicle *v : creationOrder) if (v && v->running) vehicleControl(*v);
}

void Engine::updateLocation() {
    for (Drivable *d : drivables) {
        auto &vehs = d->vehicles;
        bool stayedAhead = false;
        for (auto it = vehs.begin(); it != vehs.end();) {
            Vehicle *v = *it;
            bool leaves = v->getChangedDrivable() != nullptr || v->buffer.isEndSet;
            if (leaves) { if (stayedAhead) ++nonFifoEvents; it = vehs.erase(it); }
            else { stayedAhead = true; ++it; }
            if (v->buffer.isEndSet) {
                vehicleRemoveBuffer.insert(v);
                vehiclePool.erase(v->priority);
                cumulativeTravelTime += getCurrentTime() - v->enterTime;
                finishedVehicleCnt += 1;
                activeVehicleCount -= 1;
            }
        }
    }
    // CityFlow sorts with std::sort on distance only; exact ties are broken here
    // by creation order so that the CUDA path can state the same total order.
    std::sort(pushBuffer.begin(), pushBuffer.end(), [](const std::pair<Vehicle *, double> &a, const std::pair<Vehicle *, double> &b) {
        if (a.second != b.second) return a.second > b.second;
        return a.first->uid < b.first->uid;
    });
    for (auto &pr : pushBuffer) {
        Vehicle *v = pr.first;
        Drivable *d = v->getChangedDrivable();
        if (d != nullptr) {
            d->vehicles.push_back(v);
            v->buffer.enterLaneLinkTime = d->isLaneLink() ? (int) step : INT_MAX;
            v->buffer.isEnterLaneLinkTimeSet = true;
        }
    }
    pushBuffer.clear();
}

void Engine::updateAction() {
    for (Vehicle *v : creationOrder) {
        if (v && v->running && !vehicleRemoveBuffer.count(v)) {
            if (v->buffer.isBlockerSet && vehicleRemoveBuffer.count(v->buffer.blocker)) v->setBlocker(nullptr);
            v->update();
        }
    }
    for (Vehicle *v : vehicleRemoveBuffer) { creationOrder[v->uid] = nullptr; delete v; }
    vehicleRemoveBuffer.clear();
}

void Engine::updateLeaderAndGap() {
    for (Drivable *d : drivables) {
        Vehicle *leader = nullptr;
        for (Vehicle *v : d->vehicles) { v->updateLeaderAndGap(leader); leader = v; }
    }
}

void Engine::nextStep() {
    for (Flow &f : flows) flowStep(f);
    planRoute();
    handleWaiting();
    notifyCross();
    getAction();
    updateLocation();
    updateAction();
    updateLeaderAndGap();
    // rlTrafficLight == true: phases only change through set_tl_phase
    if (!rlTrafficLight) {
        // fixed-plan lights are not used by pytsc (config.yaml: rl_traffic_light True)
    }
    step += 1;
}

double Engine::getAverageTravelTime() const {
    double tt = cumulativeTravelTime;
    int n = finishedVehicleCnt;
    for (auto &kv : vehiclePool) { tt += getCurrentTime() - kv.second->enterTime; n++; }
    return n == 0 ? 0 : tt / n;
}

void Engine::reset(bool resetRnd) {
    for (auto &kv : vehiclePool) delete kv.second;
    vehiclePool.clear(); creationOrder.clear(); vehicleNames.clear();
    vehicleRemoveBuffer.clear(); pushBuffer.clear();
    for (Drivable *d : drivables) d->vehicles.clear();
    for (Lane *l : lanes) l->waitingBuffer.clear();
    for (auto &r : roads) r->planRouteBuffer.clear();
    for (auto &it : intersections) { it->curPhase = 0; for (Cross &c : it->crosses) c.clearNotify(); }
    for (Flow &f : flows) f.reset();
    finishedVehicleCnt = 0; cumulativeTravelTime = 0; step = 0; activeVehicleCount = 0;
    if (resetRnd) rnd.seed(seed);
}

// -- Loading --------------------------------------------------------------------
void Engine::loadRoadNet(const std::string &path) {
    Json root = readJsonFile(path);
    const Json &jint = root.at("intersections");
    const Json &jroads = root.at("roads");
    for (auto &jr : jroads.arr) {
        auto r = std::make_unique<Road>();
        r->id = jr.at("id").str;
        roadMap[r->id] = r.get();
        roads.push_back(std::move(r));
    }
    for (auto &ji : jint.arr) {
        auto it = std::make_unique<Intersection>();
        it->id = ji.at("id").str;
        interMap[it->id] = it.get();
        intersections.push_back(std::move(it));
    }
    for (size_t i = 0; i < jroads.arr.size(); ++i) {
        const Json &jr = jroads.arr[i];
        Road &r = *roads[i];
        r.startIntersection = interMap.at(jr.at("startIntersection").str);
        r.endIntersection = interMap.at(jr.at("endIntersection").str);
        for (auto &jp : jr.at("points").arr) r.points.push_back({jp.at("x").num, jp.at("y").num});
        int li = 0;
        for (auto &jl : jr.at("lanes").arr) {
            auto l = std::make_unique<Lane>();
            l->lane = true; l->road = &r; l->laneIndex = li;
            l->width = jl.at("width").num; l->maxSpeed = jl.at("maxSpeed").num;
            l->id = r.id + "_" + std::to_string(li);
            ++li;
            r.lanes.push_back(std::move(l));
        }
    }
    for (size_t i = 0; i < jint.arr.size(); ++i) {
        const Json &ji = jint.arr[i];
        Intersection &it = *intersections[i];
        it.isVirtual = ji.boolean("virtual", false);
        it.width = ji.number("width", 0);
        it.point = {ji.at("point").at("x").num, ji.at("point").at("y").num};
        for (auto &jr : ji.at("roads").arr) it.roads.push_back(roadMap.at(jr.str));
        if (it.isVirtual) continue;
        int rli = 0;
        for (auto &jrl : ji.at("roadLinks").arr) {
            auto rl = std::make_unique<RoadLink>();
            const std::string &t = jrl.at("type").str;
            rl->type = t == "go_straight" ? go_straight : (t == "turn_left" ? turn_left : turn_right);
            rl->startRoad = roadMap.at(jrl.at("startRoad").str);
            rl->endRoad = roadMap.at(jrl.at("endRoad").str);
            rl->intersection = &it; rl->index = rli++;
            for (auto &jll : jrl.at("laneLinks").arr) {
                auto ll = std::make_unique<LaneLink>();
                ll->lane = false; ll->roadLink = rl.get();
                ll->startLane = rl->startRoad->lanes.at((size_t) jll.at("startLaneIndex").num).get();
                ll->endLane = rl->endRoad->lanes.at((size_t) jll.at("endLaneIndex").num).get();
                if (const Json *jp = jll.find("points"))
                    for (auto &p : jp->arr) ll->points.push_back({p.at("x").num, p.at("y").num});
                if (ll->points.size() < 2)
                    throw std::runtime_error("lane-link without points is not supported by the oracle");
                ll->length = polylineLength(ll->points);
                ll->maxSpeed = 10000;
                ll->startLane->laneLinks.push_back(ll.get());
                it.laneLinks.push_back(ll.get());
                rl->laneLinks.push_back(std::move(ll));
            }
            it.roadLinks.push_back(std::move(rl));
        }
        const Json &tl = ji.at("trafficLight");
        for (auto &jp : tl.at("lightphases").arr) {
            LightPhase ph;
            ph.time = jp.at("time").num;
            ph.roadLinkAvailable.assign(it.roadLinks.size(), 0);
            for (auto &a : jp.at("availableRoadLinks").arr) ph.roadLinkAvailable.at((size_t) a.num) = 1;
            it.phases.push_back(std::move(ph));
        }
    }
    for (auto &it : intersections) it->initCrosses();
    for (auto &r : roads) r->initLanesPoints();
    for (auto &r : roads) for (auto &l : r->lanes) { l->index = (int) lanes.size(); lanes.push_back(l.get()); drivables.push_back(l.get()); }
    for (auto &it : intersections) for (LaneLink *ll : it->laneLinks) {
        ll->index = (int) (lanes.size() + laneLinks.size()); laneLinks.push_back(ll); drivables.push_back(ll);
    }
}

void Engine::loadFlow(const std::string &path) {
    Json root = readJsonFile(path);
    int idx = 0;
    for (auto &jf : root.arr) {
        Flow f;
        const Json &jv = jf.at("vehicle");
        VehicleInfo vi;
        vi.len = jv.number("length", vi.len); vi.width = jv.number("width", vi.width);
        vi.maxPosAcc = jv.number("maxPosAcc", vi.maxPosAcc); vi.maxNegAcc = jv.number("maxNegAcc", vi.maxNegAcc);
        vi.usualPosAcc = jv.number("usualPosAcc", vi.usualPosAcc); vi.usualNegAcc = jv.number("usualNegAcc", vi.usualNegAcc);
        vi.minGap = jv.number("minGap", vi.minGap); vi.maxSpeed = jv.number("maxSpeed", vi.maxSpeed);
        vi.headwayTime = jv.number("headwayTime", vi.headwayTime);
        vi.yieldDistance = jv.number("yieldDistance", vi.yieldDistance);
        vi.turnSpeed = jv.number("turnSpeed", vi.turnSpeed);
        f.tmpl = vi;
        for (auto &jr : jf.at("route").arr) f.route.push_back(roadMap.at(jr.str));
        f.interval = jf.number("interval", 1.0);
        f.startTime = jf.number("startTime", 0);
        f.endTime = jf.number("endTime", -1);
        f.index = idx++;
        f.reset();
        flows.push_back(std::move(f));
    }
}

void Engine::loadConfig(const std::string &cfgFile) {
    Json cfg = readJsonFile(cfgFile);
    interval = cfg.number("interval", 1.0);
    seed = (int) cfg.number("seed", 0);
    rlTrafficLight = cfg.boolean("rlTrafficLight", true);
    laneChange = cfg.boolean("laneChange", false);
    if (laneChange) throw std::runtime_error("laneChange is not restated by the oracle (pytsc default is False)");
    if (!rlTrafficLight) throw std::runtime_error("rlTrafficLight False is not restated by the oracle (pytsc default is True)");
    std::string dir = cfg.at("dir").str;
    rnd.seed(seed);
    loadRoadNet(dir + cfg.at("roadnetFile").str);
    loadFlow(dir + cfg.at("flowFile").str);
}

}  // namespace cfo

// ============================================================================
// C ABI used by oracle/engine.py (ctypes).
// ============================================================================
using cfo::Engine;
static thread_local std::string g_err;

extern "C" {

const char *cfo_last_error() { return g_err.c_str(); }

void *cfo_create(const char *config_file, int /*thread_num*/) {
    try {
        auto *e = new Engine();
        e->loadConfig(config_file);
        return e;
    } catch (std::exception &ex) { g_err = ex.what(); return nullptr; }
}
void cfo_destroy(void *h) { if (h) { auto *e = (Engine *) h; e->reset(false); delete e; } }
void cfo_next_step(void *h) { ((Engine *) h)->nextStep(); }
void cfo_next_steps(void *h, int n) { for (int i = 0; i < n; ++i) ((Engine *) h)->nextStep(); }
double cfo_get_current_time(void *h) { return ((Engine *) h)->getCurrentTime(); }
void cfo_reset(void *h, int resetRnd) { ((Engine *) h)->reset(resetRnd != 0); }
int cfo_set_tl_phase(void *h, const char *id, int phase) {
    auto *e = (Engine *) h;
    auto it = e->interMap.find(id);
    if (it == e->interMap.end() || phase < 0 || phase >= (int) it->second->phases.size()) return -1;
    it->second->curPhase = phase;
    return 0;
}
int cfo_set_tl_phase_idx(void *h, int inter, int phase) {
    auto *e = (Engine *) h;
    if (inter < 0 || inter >= (int) e->intersections.size()) return -1;
    if (phase < 0 || phase >= (int) e->intersections[inter]->phases.size()) return -1;
    e->intersections[inter]->curPhase = phase;
    return 0;
}
int cfo_get_vehicle_count(void *h) { return ((Engine *) h)->activeVehicleCount; }
int cfo_get_finished_count(void *h) { return ((Engine *) h)->finishedVehicleCnt; }
int cfo_get_created_count(void *h) { return (int) ((Engine *) h)->creationOrder.size(); }
long long cfo_get_non_fifo_events(void *h) { return ((Engine *) h)->nonFifoEvents; }
double cfo_get_average_travel_time(void *h) { return ((Engine *) h)->getAverageTravelTime(); }

int cfo_num_lanes(void *h) { return (int) ((Engine *) h)->lanes.size(); }
int cfo_num_lanelinks(void *h) { return (int) ((Engine *) h)->laneLinks.size(); }
int cfo_num_intersections(void *h) { return (int) ((Engine *) h)->intersections.size(); }
const char *cfo_lane_id(void *h, int i) { return ((Engine *) h)->lanes[i]->id.c_str(); }
const char *cfo_intersection_id(void *h, int i) { return ((Engine *) h)->intersections[i]->id.c_str(); }
double cfo_drivable_length(void *h, int d) { return ((Engine *) h)->drivables[d]->length; }
int cfo_num_crosses(void *h, int inter) { return (int) ((Engine *) h)->intersections[inter]->crosses.size(); }
// per cross: lane-link drivable indices and distances (for checking the scenario compiler)
int cfo_get_crosses(void *h, int inter, int *ll0, int *ll1, double *d0, double *d1) {
    auto &cs = ((Engine *) h)->intersections[inter]->crosses;
    for (size_t i = 0; i < cs.size(); ++i) {
        ll0[i] = cs[i].laneLinks[0]->index; ll1[i] = cs[i].laneLinks[1]->index;
        d0[i] = cs[i].distanceOnLane[0]; d1[i] = cs[i].distanceOnLane[1];
    }
    return (int) cs.size();
}

// get_lane_waiting_vehicle_count / get_lane_vehicles: per lane in roadnet order
void cfo_lane_counts(void *h, int *n_vehicles, int *n_waiting) {
    auto *e = (Engine *) h;
    for (size_t i = 0; i < e->lanes.size(); ++i) {
        int w = 0;
        for (auto *v : e->lanes[i]->vehicles) if (v->info.speed < 0.1) ++w;
        n_vehicles[i] = (int) e->lanes[i]->vehicles.size();
        n_waiting[i] = w;
    }
}
// concatenated vehicle uids of every lane, front to back; returns total
int cfo_lane_vehicles(void *h, int *uids, int cap) {
    auto *e = (Engine *) h;
    int k = 0;
    for (auto *l : e->lanes) for (auto *v : l->vehicles) { if (k < cap) uids[k] = v->uid; ++k; }
    return k;
}
// every running vehicle, drivable-major / list order; returns total
int cfo_running_vehicles(void *h, int *uid, int *drivable, double *dist, double *speed,
                         int *leader_uid, double *gap, int *blocker_uid, int *enter_ll_time, int *priority, int cap) {
    auto *e = (Engine *) h;
    int k = 0;
    for (auto *d : e->drivables) for (auto *v : d->vehicles) {
        if (k < cap) {
            if (uid) uid[k] = v->uid;
            if (drivable) drivable[k] = d->index;
            if (dist) dist[k] = v->dis;
            if (speed) speed[k] = v->info.speed;
            if (leader_uid) leader_uid[k] = v->leader ? v->leader->uid : -1;
            if (gap) gap[k] = v->leader ? v->gap : 0.0;
            if (blocker_uid) blocker_uid[k] = v->blocker ? v->blocker->uid : -1;
            if (enter_ll_time) enter_ll_time[k] = v->enterLaneLinkTime;
            if (priority) priority[k] = v->priority;
        }
        ++k;
    }
    return k;
}
const char *cfo_vehicle_name(void *h, int uid) {
    auto *e = (Engine *) h;
    if (uid < 0 || uid >= (int) e->vehicleNames.size()) return "";
    return e->vehicleNames[uid].c_str();
}
// running flag, distance, speed, drivable index for one vehicle (get_vehicle_info)
int cfo_vehicle_info(void *h, int uid, double *dist, double *speed, int *drivable) {
    auto *e = (Engine *) h;
    if (uid < 0 || uid >= (int) e->creationOrder.size()) return -1;
    cfo::Vehicle *v = e->creationOrder[uid];
    if (!v) return -1;
    if (!v->running) return 0;
    *dist = v->dis; *speed = v->info.speed; *drivable = v->drivable->index;
    return 1;
}
// waiting-buffer sizes per lane
void cfo_waiting_buffer_sizes(void *h, int *out) {
    auto *e = (Engine *) h;
    for (size_t i = 0; i < e->lanes.size(); ++i) out[i] = (int) e->lanes[i]->waitingBuffer.size();
}

}  // extern "C"
