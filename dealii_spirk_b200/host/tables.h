// Butcher-tableau loaders (ref main.cc:599-656) and a minimal ConvergenceTable / JSON reader.
#pragma once
#include <cstdlib>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <map>
#include <sstream>
#include <string>
#include <vector>

#include "vector.h"

namespace spirk_host
{
  struct FullMatrix
  {
    unsigned int        rows = 0, cols = 0;
    std::vector<double> v;
    FullMatrix() = default;
    FullMatrix(unsigned int m, unsigned int n)
      : rows(m)
      , cols(n)
      , v((size_t)m * n, 0.0)
    {}
    double       &operator()(unsigned int i, unsigned int j) { return v[(size_t)i * cols + j]; }
    const double &operator()(unsigned int i, unsigned int j) const { return v[(size_t)i * cols + j]; }
    const double *operator[](unsigned int i) const { return &v[(size_t)i * cols]; }
    unsigned int  m() const { return rows; }
    unsigned int  n() const { return cols; }
  };

  // directory / packed file the tables come from; set by the driver (default: $SPIRK_TABLES)
  inline std::string &tables_location()
  {
    static std::string loc = std::getenv("SPIRK_TABLES") ? std::getenv("SPIRK_TABLES") : "";
    return loc;
  }

  namespace internal
  {
    // reference text format: "m n v0 v1 ..." (whitespace separated), file `<label><q>.txt` looked up in
    // the working directory, then in ../ (ref main.cc:604-610); then the packed file
    // dealii_spirk_b200/tables/butcher_tables.txt (one record per line: label q m n values...).
    inline bool read_table(const unsigned int n_stages, const std::string &label, unsigned int &m, unsigned int &n,
                           std::vector<double> &values)
    {
      const std::string file_name = label + std::to_string(n_stages) + ".txt";
      std::ifstream     fin(file_name);
      if (fin.fail())
        fin.open("../" + file_name);
      if (!fin.fail())
        {
          fin >> m >> n;
          values.resize((size_t)m * n);
          for (auto &x : values)
            fin >> x;
          return !fin.fail();
        }
      std::vector<std::string> candidates;
      if (!tables_location().empty())
        {
          candidates.push_back(tables_location());
          candidates.push_back(tables_location() + "/butcher_tables.txt");
        }
      candidates.push_back("butcher_tables.txt");
      candidates.push_back("tables/butcher_tables.txt");
      candidates.push_back("dealii_spirk_b200/tables/butcher_tables.txt");
      for (const auto &c : candidates)
        {
          std::ifstream pk(c);
          if (pk.fail())
            continue;
          std::string line;
          while (std::getline(pk, line))
            {
              if (line.empty() || line[0] == '#')
                continue;
              std::istringstream is(line);
              std::string        lab;
              unsigned int       q;
              is >> lab >> q;
              if (lab != label || q != n_stages)
                continue;
              is >> m >> n;
              values.resize((size_t)m * n);
              for (auto &x : values)
                is >> x;
              return !is.fail();
            }
        }
      return false;
    }
  } // namespace internal

  inline FullMatrix load_matrix_from_file(const unsigned int n_stages, const std::string label)
  {
    unsigned int        m, n;
    std::vector<double> v;
    if (!internal::read_table(n_stages, label, m, n, v))
      throw Error("File with the name " + label + std::to_string(n_stages) + ".txt could not be found!");
    if (m != n_stages || n != n_stages)
      throw Error("table " + label + ": dimension mismatch");
    FullMatrix result(n_stages, n_stages);
    result.v = v;
    return result;
  }

  inline std::vector<double> load_vector_from_file(const unsigned int n_stages, const std::string label)
  {
    unsigned int        m, n;
    std::vector<double> v;
    if (!internal::read_table(n_stages, label, m, n, v))
      throw Error("File with the name " + label + std::to_string(n_stages) + ".txt could not be found!");
    if (m != 1 || n != n_stages)
      throw Error("table " + label + ": dimension mismatch");
    return v;
  }

  // the columns the reference's ConvergenceTable carries (main.cc:689-719, 3360-3368, 3387-3398)
  class ConvergenceTable
  {
  public:
    void add_value(const std::string &key, double value)
    {
      if (!columns.count(key))
        order.push_back(key);
      columns[key].push_back(value);
    }
    void set_scientific(const std::string &key, bool s) { scientific[key] = s; }
    void write_text(std::ostream &out) const
    {
      for (const auto &k : order)
        out << std::setw(15) << k << " ";
      out << std::endl;
      size_t rows = 0;
      for (const auto &k : order)
        rows = std::max(rows, columns.at(k).size());
      for (size_t r = 0; r < rows; ++r)
        {
          for (const auto &k : order)
            {
              const auto &c = columns.at(k);
              std::ostringstream s;
              if (r < c.size())
                {
                  if (scientific.count(k) && scientific.at(k))
                    s << std::scientific << std::setprecision(4) << c[r];
                  else
                    s << c[r];
                }
              out << std::setw(15) << s.str() << " ";
            }
          out << std::endl;
        }
    }
    std::vector<std::string>                   order;
    std::map<std::string, std::vector<double>> columns;
    std::map<std::string, bool>                scientific;
  };

  // flat JSON object reader: {"Key": value, ...} with string / number / bool values, which is all the
  // reference's parameter files contain (json/*.json, scripts/default.json)
  inline std::map<std::string, std::string> parse_flat_json(const std::string &text)
  {
    std::map<std::string, std::string> out;
    size_t                             i = 0;
    auto skip = [&]() {
      while (i < text.size() && std::isspace((unsigned char)text[i]))
        ++i;
    };
    auto str = [&]() {
      std::string s;
      ++i;
      while (i < text.size() && text[i] != '"')
        {
          if (text[i] == '\\' && i + 1 < text.size())
            ++i;
          s += text[i++];
        }
      ++i;
      return s;
    };
    skip();
    if (i >= text.size() || text[i] != '{')
      throw Error("parameter file: expected '{'");
    ++i;
    while (true)
      {
        skip();
        if (i < text.size() && text[i] == '}')
          break;
        if (i >= text.size() || text[i] != '"')
          throw Error("parameter file: expected a key string");
        const std::string key = str();
        skip();
        if (i >= text.size() || text[i] != ':')
          throw Error("parameter file: expected ':' after key " + key);
        ++i;
        skip();
        std::string val;
        if (text[i] == '"')
          val = str();
        else
          while (i < text.size() && text[i] != ',' && text[i] != '}' && !std::isspace((unsigned char)text[i]))
            val += text[i++];
        out[key] = val;
        skip();
        if (i < text.size() && text[i] == ',')
          ++i;
      }
    return out;
  }
} // namespace spirk_host
