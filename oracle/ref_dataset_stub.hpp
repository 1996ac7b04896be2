// ref_dataset_stub.hpp -- TEST INFRASTRUCTURE.  An in-memory stand-in for the reference's GDCM-backed
// mqi::dataset (moqui/base/mqi_dataset.hpp, which cannot be compiled here: GDCM is an un-vendored third-party
// dependency).  It offers the interface the GDCM-free reference headers above it use -- keyword
// lookups into typed vectors and named sub-sequences -- so that treatment_machine_ion / treatment_machine_pbs
// (beam model, spot -> beamlet, histories per spot, range shifter, aperture, beam frame) compile UNMODIFIED from
// /root/reference and run on a plan built in memory by oracle/ref_tps_kat.cpp.  build_ref.sh copies this file to a
// scratch include directory as moqui/base/mqi_dataset.hpp; it is written from scratch, not derived from the
// reference file it shadows.
#ifndef MQI_DATASET_H
#define MQI_DATASET_H
#include <algorithm>
#include <cstdlib>
#include <map>
#include <string>
#include <vector>

namespace mqi
{
typedef enum { RTPLAN, IONPLAN, RTRECORD, IONRECORD, RTIMAGE, RTSTRUCT, RTDOSE, UNKNOWN_MOD } modality_type;

struct beam_id_type {
    enum { NUM, STR } type;
    union {
        int         number;
        const char* name;
    };
};

// a sequence is addressed by its name in the stub
struct seq_tag {
    std::string name;
};

static const std::map<const modality_type, const std::map<const std::string, const seq_tag>> seqtags_per_modality = {
    { IONPLAN,
      { { "beam", { "beam" } }, { "snout", { "snout" } }, { "rs", { "rs" } }, { "rsss", { "rsss" } }, { "blk", { "blk" } },
        { "comp", { "comp" } }, { "ctrl", { "ctrl" } } } },
    { IONRECORD,
      { { "beam", { "beam" } }, { "snout", { "snout" } }, { "rs", { "rs" } }, { "rsss", { "rsss" } }, { "blk", { "blk" } },
        { "comp", { "comp" } }, { "ctrl", { "ctrl" } } } }
};

class dataset
{
public:
    std::map<std::string, std::vector<std::string>>    values;      // keyword -> values as text
    std::map<std::string, std::vector<const dataset*>> sequences;   // sequence name -> items

    dataset& set(const std::string& k, std::vector<std::string> v) { values[k] = std::move(v); return *this; }
    dataset& add(const std::string& seq, const dataset* item) { sequences[seq].push_back(item); return *this; }

    std::vector<const dataset*>
    operator()(const seq_tag& t) const {
        auto it = sequences.find(t.name);
        return it == sequences.end() ? std::vector<const dataset*>() : it->second;
    }
    std::vector<const dataset*>
    operator()(const char* s) const { return (*this)(seq_tag{ s }); }

    void get_values(const char* k, std::vector<std::string>& out) const {
        out.clear();
        auto it = values.find(k);
        if (it != values.end()) out = it->second;
    }
    void get_values(const char* k, std::vector<float>& out) const {
        out.clear();
        auto it = values.find(k);
        if (it != values.end()) for (const auto& s : it->second) out.push_back((float) std::atof(s.c_str()));
    }
    void get_values(const char* k, std::vector<int>& out) const {
        out.clear();
        auto it = values.find(k);
        if (it != values.end()) for (const auto& s : it->second) out.push_back(std::atoi(s.c_str()));
    }
    void get_values(const char* k, std::vector<double>& out) const {
        out.clear();
        auto it = values.find(k);
        if (it != values.end()) for (const auto& s : it->second) out.push_back(std::atof(s.c_str()));
    }
    void dump() const {}
};
}   // namespace mqi
#endif
