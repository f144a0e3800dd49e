// main.cpp -- the SM_${NS}x${NT} executable on B200s.
//
// Drop-in for the reference's program (src/main.cpp there): the same ten parameters are read from
// stdin in the same order with the same prompts on stderr, the same banner and result lines go
// to stdout, and the same 2D_U1_{Nx}x{Nt}_m0{m0}_SimData.txt log is written.  What differs is
// underneath: no mpirun -- after the prompts the program forks ranks_x*ranks_t - 1 worker
// processes itself, one per GPU, and every trajectory runs device-resident in libschwinger_b200.
#include <chrono>
#include <ctime>
#include <fstream>
#include <functional>
#include <iomanip>
#include <sstream>
#include <string>
#include <vector>

#include "b200_runtime.h"
#include "hmc.h"
#include "mpi_setup.h"

namespace {

struct RunParams {
    int ranks_x = 1, ranks_t = 1;
    double m0 = 0;
    int md_steps = 0;
    double trajectory_length = 0, beta = 0;
    int ntherm = 0, nmeas = 0, nsteps = 0, saveconf = 0;
};

// prompt text (reference: src/main.cpp:37-56) paired with the field it fills, in input order
bool read_params(RunParams& p) {
    const std::vector<std::pair<const char*, std::function<void()>>> questions = {
        {"ranks_x: ", [&] { std::cin >> p.ranks_x; }},
        {"ranks_t: ", [&] { std::cin >> p.ranks_t; }},
        {"m0: ", [&] { std::cin >> p.m0; }},
        {"Molecular dynamics steps: ", [&] { std::cin >> p.md_steps; }},
        {"Trajectory length: ", [&] { std::cin >> p.trajectory_length; }},
        {"beta: ", [&] { std::cin >> p.beta; }},
        {"Thermalization: ", [&] { std::cin >> p.ntherm; }},
        {"Measurements: ", [&] { std::cin >> p.nmeas; }},
        {"Step (sweeps between measurements): ", [&] { std::cin >> p.nsteps; }},
        {"Save configurations yes/no (1 or 0): ", [&] { std::cin >> p.saveconf; }},
    };
    for (const char* line : {"  -----------------------------", "|  Two-flavor Schwinger model   |",
                             "| Hybrid Monte Carlo simulation |", "  -----------------------------"})
        std::cerr << line << std::endl;
    std::cerr << "Nx " << LV::Nx << " Nt " << LV::Nt << std::endl;
    for (const auto& q : questions) {
        std::cerr << q.first << std::endl;
        q.second();
    }
    std::cerr << std::endl;
    return bool(std::cin) && p.ranks_x >= 1 && p.ranks_t >= 1;
}

std::string timestamp_now() {
    const std::time_t t = std::chrono::system_clock::to_time_t(std::chrono::system_clock::now());
    std::ostringstream s;
    s << std::put_time(std::localtime(&t), "%Y-%m-%d %H:%M:%S");
    return s.str();
}

std::string simdata_name(double m0) {
    std::ostringstream s;
    s << "2D_U1_" << LV::Nx << "x" << LV::Nt << "_m0" << std::setprecision(17) << m0 << "_SimData.txt";
    return s.str();
}

// a labelled row of right-aligned columns, the layout of the reference's log (src/main.cpp:107-125)
class Row {
public:
    explicit Row(std::ostream& o) : out(o) {}
    template <typename T>
    Row& col(int width, const T& v) {
        out << std::setw(width) << std::setprecision(17) << v;
        return *this;
    }
    void end() { out << "\n"; }

private:
    std::ostream& out;
};

void write_simdata_header(const std::string& file, const RunParams& p, const std::string& when, const char* host) {
    std::ofstream f(file);
    f << "#Date and time\n" << when << "\n#Host\n" << (host ? host : "unknown") << "\n";
    f << "#Nx      #Nt\n";
    Row(f).col(10, LV::Nx).col(10, LV::Nt).end();
    f << "#ranks_x     #ranks_t     #ranks\n";
    Row(f).col(15, mpi::ranks_x).col(15, mpi::ranks_t).col(15, mpi::size).end();
    f << "#beta                        #Ntherm     #Nmeas     #Nsteps\n";
    Row(f).col(30, p.beta).col(11, p.ntherm).col(11, p.nmeas).col(11, p.nsteps).end();
    f << "#trajectory_length     #MD_steps\n";
    Row(f).col(30, p.trajectory_length).col(30, p.md_steps).end();
    f << "#CG max iterations     #CG relative tolerance\n";
    Row(f).col(30, CG::max_iter).col(30, CG::tol).end();
    f << "#m0\n";
    Row(f).col(30, p.m0).end();
}

void append_simdata_results(const std::string& file, HMC& hmc, const RunParams& p, double seconds) {
    std::ofstream f(file, std::ios::app);
    f << "#Ep                           #dEp\n";
    Row(f).col(30, hmc.getEp()).col(30, hmc.getdEp()).end();
    f << "#gS                           #dgS\n";
    Row(f).col(30, hmc.getgS()).col(30, hmc.getdgS()).end();
    f << "#Acceptance rate\n";
    Row(f).col(30, hmc.getacceptance_rate(p.nmeas + p.nsteps * (p.nmeas - 1))).end();
    f << "#Execution time\n";
    f << std::setw(30) << std::setprecision(17) << seconds;   // no trailing newline, as in the reference
}

void print_banner(const RunParams& p, const std::string& when, const char* host) {
    const std::string bar(70, '*');
    std::ostream& o = std::cout;
    o << bar << std::endl;
    o << "*                              PARAMETERS" << std::endl;
    o << "* Nx = " << LV::Nx << ", Nt = " << LV::Nt << std::endl;
    o << "* m0 = " << p.m0 << ", kappa = " << 1 / (2 * (p.m0 + 2)) << std::endl;
    o << "* beta = " << p.beta << std::endl;
    o << "* Thermalization confs = " << p.ntherm << std::endl;
    o << "* Measurement confs = " << p.nmeas << std::endl;
    o << "* Decorrelation steps (confs dropped between measurements) = " << p.nsteps << std::endl;
    o << "* Trajectory length = " << p.trajectory_length << ", Leapfrog steps = " << p.md_steps
      << ", Integration step = " << p.trajectory_length / p.md_steps << std::endl;
    o << "* CG max iterations = " << CG::max_iter << ", CG tolerance = " << CG::tol << std::endl;
    o << "* Number of ranks on x = " << mpi::ranks_x << ", Number of ranks on t = " << mpi::ranks_t << std::endl;
    o << "* Total number of MPI ranks = " << mpi::size << std::endl;
    o << "* Each rank has " << mpi::maxSize << " lattice sites" << std::endl;
    o << "* Host: " << (host ? host : "unknown") << std::endl;
    o << "* Start time: " << when << std::endl;
    o << bar << std::endl;
}

// B200 work counters go to their own file so that _SimData.txt stays byte-compatible
void write_work_counters(const HMC& hmc, double seconds) {
    std::ostringstream name;
    name << "2D_U1_" << LV::Nx << "x" << LV::Nt << "_b200.json";
    std::ofstream js(name.str());
    const double site_updates = (double)hmc.dd_applications() * LV::Ntot;
    js << std::setprecision(12) << "{\"trajectories\": " << hmc.trajectories() << ", \"seconds\": " << seconds
       << ", \"traj_per_s\": " << hmc.trajectories() / seconds << ", \"device_seconds\": " << hmc.device_seconds()
       << ", \"dd_applications\": " << hmc.dd_applications() << ", \"dd_site_updates_per_s\": "
       << (hmc.device_seconds() > 0 ? site_updates / hmc.device_seconds() : 0.0) << ", \"gpus\": " << mpi::size << "}\n";
}

}  // namespace

int main() {
    CG::max_iter = 10000;   // the reference's settings (src/main.cpp:26-27)
    CG::tol = 1e-10;

    RunParams p;
    if (!read_params(p)) {
        std::cerr << "could not read the ten run parameters from stdin" << std::endl;
        return 1;
    }
    mpi::ranks_x = p.ranks_x;
    mpi::ranks_t = p.ranks_t;

    // one process per GPU; the workers inherit every parameter through fork (the reference broadcasts them)
    b200::spawn_ranks(mpi::ranks_x * mpi::ranks_t);
    srand((mpi::rank + 1) * time(0));

    initializeMPI();            // tile widths, neighbour ranks, GPU context (+ NCCL)
    allocate_lattice_arrays();
    periodic_boundary();

    GaugeConf GConf;
    const bool root = (mpi::rank == 0);
    const std::string when = timestamp_now();
    const char* host = std::getenv("HOSTNAME");
    const std::string log = simdata_name(p.m0);
    if (root) {
        write_simdata_header(log, p, when, host);
        print_banner(p, when, host);
    }

    const char* start = std::getenv("SM_START_CONF");   // optional: resume from a .ctxt file
    if (start) GConf.readBinary(start);
    HMC hmc(GConf, p.md_steps, p.trajectory_length, p.ntherm, p.nmeas, p.nsteps, p.beta, LV::Nx, LV::Nt, p.m0, p.saveconf);
    hmc.set_start_from_conf(start != nullptr);

    const double t0 = b200::wtime();
    hmc.HMC_algorithm();
    const double seconds = b200::wtime() - t0;

    if (root) {
        std::cout << "Average plaquette value / volume: Ep = " << hmc.getEp() << " dEp = " << hmc.getdEp() << std::endl;
        std::cout << "Average gauge action / volume: gS = " << hmc.getgS() << " dgS = " << hmc.getdgS() << std::endl;
        // the reference prints this line with Nmeas + Nsteps*Nmeas in the denominator (src/main.cpp:159)
        std::cout << "Acceptance rate: " << hmc.getacceptance_rate(p.nmeas + p.nsteps * p.nmeas) << std::endl;
        std::cout << "Execution time = " << seconds << " s" << std::endl;
        std::cout << "-------------------------------" << std::endl;
        append_simdata_results(log, hmc, p, seconds);
        write_work_counters(hmc, seconds);
    }

    free_lattice_arrays();
    b200::shutdown();
    return 0;
}
